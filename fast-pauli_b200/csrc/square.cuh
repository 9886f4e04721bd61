// K8: SummedPauliOp::square() (SPO:197-268): A_k -> A_k^2, coefficient of output string c in operator k
//
//     coeffs_sq(c, k) = sum over ordered pairs (a, b) with P_a P_b = phase(a, b) P_c of phase(a, b) h(a, k) h(b, k)
//
// The reference materialises the sparse tensor T_caij with S^2 hash-map insertions and contracts it on the host.
// Here one CTA owns one output string c.  For every input string a the partner is determined, b = the string with
// masks (x_a ^ x_c, z_a ^ z_c), found by one probe sequence of an open-addressing table of the (duplicate-merged)
// input strings, so the pair list of c is enumerated on the fly (S probes per output string, no S^2 tensor).  The
// phase follows from the masks alone: with P(x, z) = i^{popc(x & z)} X^x Z^z,
//     P_a P_b = i^{nY_a + nY_b - nY_c + 2 popc(z_a & x_b)} P_c.
// Found pairs are queued in thread order (ballot + prefix: the summation order is fixed, results are reproducible)
// and contracted cooperatively: lanes run along the operator axis k (coalesced rows of h), pair slots along the rest
// of the CTA, one complex accumulator per (thread, 128-wide k chunk).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace fpk
{

struct SqEntry // open-addressing table entry, 32 bytes
{
    uint64_t x, z;
    uint32_t idx; // input string index, 0xffffffff = empty
    uint32_t ny;  // popc(x & z) mod 4
    uint64_t pad;
};

constexpr int kSqThreads = 128;
constexpr int kSqMaxKChunks = 8; // operators per launch <= 8 * 128

__device__ __forceinline__ uint32_t sq_hash(uint64_t x, uint64_t z)
{
    uint64_t h = (x * 0x9E3779B97F4A7C15ull) ^ (z * 0xC2B2AE3D27D4EB4Full);
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return static_cast<uint32_t>(h);
}

template <typename T>
__global__ void __launch_bounds__(kSqThreads)
    sop_square_kernel(uint32_t S, uint64_t const *__restrict__ xs, uint64_t const *__restrict__ zs,
                      SqEntry const *__restrict__ table, uint32_t table_mask, Cx<T> const *__restrict__ h /* [S][Kp] */,
                      uint32_t Kp /* row pitch */, uint32_t K /* operators of this launch */,
                      uint64_t const *__restrict__ xq, uint64_t const *__restrict__ zq, Cx<T> *__restrict__ out /* [n_sq][Kp] */)
{
    __shared__ uint32_t q_a[kSqThreads], q_b[kSqThreads], q_ph[kSqThreads];
    __shared__ uint32_t warp_cnt[kSqThreads / 32];
    __shared__ Cx<T> red[kSqThreads];

    uint32_t const tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint64_t const c = blockIdx.x;
    uint64_t const xc = xq[c], zc = zq[c];
    uint32_t const nyc = static_cast<uint32_t>(__popcll(xc & zc)) & 3u;

    // contraction geometry: KL lanes along k, PL pair slots
    uint32_t KL = 1;
    while (KL < K && KL < kSqThreads)
        KL <<= 1;
    uint32_t const PL = kSqThreads / KL;
    uint32_t const kl = tid % KL, slot = tid / KL;
    uint32_t const n_chunks = (K + KL - 1) / KL;

    Cx<T> acc[kSqMaxKChunks];
#pragma unroll
    for (int i = 0; i < kSqMaxKChunks; ++i)
        acc[i] = Cx<T>{0, 0};

    for (uint32_t a0 = 0; a0 < S; a0 += kSqThreads)
    {
        // ---- every thread looks up the partner of one input string
        uint32_t const a = a0 + tid;
        bool found = false;
        uint32_t b = 0, ph = 0;
        if (a < S)
        {
            uint64_t const xa = xs[a], za = zs[a];
            uint64_t const xb = xa ^ xc, zb = za ^ zc;
            uint32_t slot_i = sq_hash(xb, zb) & table_mask;
            for (;;)
            {
                SqEntry const e = table[slot_i];
                if (e.idx == 0xffffffffu)
                    break;
                if (e.x == xb && e.z == zb)
                {
                    found = true;
                    b = e.idx;
                    uint32_t const nya = static_cast<uint32_t>(__popcll(xa & za));
                    ph = (nya + e.ny + 4u - nyc + 2u * (static_cast<uint32_t>(__popcll(za & xb)) & 1u)) & 3u;
                    break;
                }
                slot_i = (slot_i + 1) & table_mask;
            }
        }
        // ---- queue the found pairs in thread order
        uint32_t const ballot = __ballot_sync(0xffffffffu, found);
        if (lane == 0)
            warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        uint32_t base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kSqThreads / 32; ++w)
        {
            if (w < static_cast<int>(warp))
                base += warp_cnt[w];
            total += warp_cnt[w];
        }
        if (found)
        {
            uint32_t const pos = base + __popc(ballot & ((1u << lane) - 1u));
            q_a[pos] = a;
            q_b[pos] = b;
            q_ph[pos] = ph;
        }
        __syncthreads();
        // ---- contract: acc(k) += i^ph h(a, k) h(b, k)
        for (uint32_t p = slot; p < total; p += PL)
        {
            Cx<T> const *ha = h + static_cast<uint64_t>(q_a[p]) * Kp;
            Cx<T> const *hb = h + static_cast<uint64_t>(q_b[p]) * Kp;
            uint32_t const phs = q_ph[p];
#pragma unroll
            for (int i = 0; i < kSqMaxKChunks; ++i)
            {
                uint32_t const k = kl + i * KL;
                if (i < static_cast<int>(n_chunks) && k < K)
                {
                    Cx<T> const u = ha[k], v = hb[k];
                    T const pr = u.re * v.re - u.im * v.im, pi = u.re * v.im + u.im * v.re;
                    // multiply by i^ph: (pr, pi), (-pi, pr), (-pr, -pi), (pi, -pr)
                    T const rr = (phs & 1u) ? ((phs & 2u) ? pi : -pi) : ((phs & 2u) ? -pr : pr);
                    T const ii = (phs & 1u) ? ((phs & 2u) ? -pr : pr) : ((phs & 2u) ? -pi : pi);
                    acc[i].re += rr;
                    acc[i].im += ii;
                }
            }
        }
        __syncthreads(); // the queue is refilled by the next batch
    }

    // ---- fold the pair slots that share an operator lane, fixed order
#pragma unroll
    for (int i = 0; i < kSqMaxKChunks; ++i)
    {
        if (i >= static_cast<int>(n_chunks))
            break;
        red[tid] = acc[i];
        __syncthreads();
        if (slot == 0)
        {
            Cx<T> sum = red[kl];
            for (uint32_t s = 1; s < PL; ++s)
            {
                sum.re += red[s * KL + kl].re;
                sum.im += red[s * KL + kl].im;
            }
            uint32_t const k = kl + i * KL;
            if (k < K)
                out[c * Kp + k] = sum;
        }
        __syncthreads();
    }
}

} // namespace fpk
