// K4c: SummedPauliOp::expectation_value first stage (SPO:573-577), E(s, t) = <psi_t| P_s |psi_t> for every string,
// for registers of 9..12 qubits (complex64: a column pair per CTA in packed FP32; complex128: one column, FP64).  Same plan as sop_expval_tile_kernel (coset.cuh) -- the whole state
// column pair lives in shared memory, warps take the x-mask chunks (<= MS strings sharing a gather), rows are
// visited as unordered pairs {i, i^x} so q = conj(psi_i) psi_{i^x} is formed once per pair -- with the per-element
// integer work removed:
//
//   * the tile is stored planar per pair, (re0, re1, im0, im1): q of both columns is two packed FP32 ops;
//   * a lane visits pair indices p = lane + 32 j; the sign (-1)^{popc(i & z)} factors into a lane part (applied
//     once at the very end), a block part (j >> 3, one flip per 8 pairs) and a part that depends on j & 7 only
//     through three bits of z: q of 8 consecutive j is kept in registers and a warp-uniform 8-way switch adds it
//     up with compile-time signs -- one FADD2 per (string, pair) instead of popc / xor / flip / add per column.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "coset.cuh"

namespace fpk
{

__device__ __forceinline__ float2 f2neg(float2 a)
{
    return make_float2(-a.x, -a.y);
}
__device__ __forceinline__ float2 f2flip(float2 a, uint32_t odd)
{
    uint32_t const s = odd << 31;
    return make_float2(__uint_as_float(__float_as_uint(a.x) ^ s), __uint_as_float(__float_as_uint(a.y) ^ s));
}

// Arithmetic of one shared-memory row: complex64 = a column PAIR as packed float2 lanes, complex128 = one column.
struct EtF32
{
    using T = float;
    using V = float2;  // one value per column of the pair
    using Row = float4; // planar (re0, re1, im0, im1)
    static constexpr int COLS = 2;
    static __device__ __forceinline__ Row to_row(float4 a) { return make_float4(a.x, a.z, a.y, a.w); }
    static __device__ __forceinline__ V re(Row a) { return make_float2(a.x, a.y); }
    static __device__ __forceinline__ V im(Row a) { return make_float2(a.z, a.w); }
    static __device__ __forceinline__ V zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ V add(V a, V b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ V mul(V a, V b) { return __fmul2_rn(a, b); }
    static __device__ __forceinline__ V fma(V a, V b, V c) { return __ffma2_rn(a, b, c); }
    static __device__ __forceinline__ V neg(V a) { return f2neg(a); }
    static __device__ __forceinline__ V flip(V a, uint32_t odd) { return f2flip(a, odd); }
    static __device__ __forceinline__ V shfl_add(V a, int off)
    {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
        return a;
    }
    static __device__ __forceinline__ void store(T *E, uint64_t s, uint64_t B, uint64_t v, V val, bool accumulate)
    {
        float2 *p = reinterpret_cast<float2 *>(E + s * B + v * 2);
        if (accumulate)
        {
            float2 const o = *p;
            val.x += o.x;
            val.y += o.y;
        }
        *p = val;
    }
};
struct EtF64
{
    using T = double;
    using V = double;
    using Row = double2; // (re, im)
    static constexpr int COLS = 1;
    static __device__ __forceinline__ Row to_row(double2 a) { return a; }
    static __device__ __forceinline__ V re(Row a) { return a.x; }
    static __device__ __forceinline__ V im(Row a) { return a.y; }
    static __device__ __forceinline__ V zero() { return 0.0; }
    static __device__ __forceinline__ V add(V a, V b) { return a + b; }
    static __device__ __forceinline__ V mul(V a, V b) { return a * b; }
    static __device__ __forceinline__ V fma(V a, V b, V c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ V neg(V a) { return -a; }
    static __device__ __forceinline__ V flip(V a, uint32_t odd) { return flip_sign(a, odd); }
    static __device__ __forceinline__ V shfl_add(V a, int off) { return a + __shfl_xor_sync(0xffffffffu, a, off); }
    static __device__ __forceinline__ void store(T *E, uint64_t s, uint64_t B, uint64_t v, V val, bool accumulate)
    {
        E[s * B + v] = accumulate ? E[s * B + v] + val : val;
    }
};

// sum_k (-1)^{popc(k & K)} q[k], k < 8, K compile time (FADD2 / DADD negate an operand for free)
template <typename P, int K> __device__ __forceinline__ typename P::V signed_sum8(typename P::V const (&q)[8])
{
    typename P::V p = q[0];
#pragma unroll
    for (int k = 1; k < 8; ++k)
        p = P::add(p, (__builtin_popcount(k & K) & 1) ? P::neg(q[k]) : q[k]);
    return p;
}

template <typename P> __device__ __forceinline__ typename P::V signed_sum8_dyn(uint32_t zk, typename P::V const (&q)[8])
{
    switch (zk) // warp-uniform
    {
    case 0:
        return signed_sum8<P, 0>(q);
    case 1:
        return signed_sum8<P, 1>(q);
    case 2:
        return signed_sum8<P, 2>(q);
    case 3:
        return signed_sum8<P, 3>(q);
    case 4:
        return signed_sum8<P, 4>(q);
    case 5:
        return signed_sum8<P, 5>(q);
    case 6:
        return signed_sum8<P, 6>(q);
    default:
        return signed_sum8<P, 7>(q);
    }
}

// All blocks of one chunk.  ADDR: how the 8 pairs of a block are laid out behind the block's first row --
// 0: rows 32 apart (512-byte immediates; diagonal chunks and x-masks whose top bit is >= 8), 1: rows 64 apart
// (top bit < 5: the pair bit sits among the lane bits), 2: the pair bit sits among the three in-block bits
// (top bit 5..7): offsets from a register table.  DIAG: x = 0, q = |psi|^2.  ODD: some string of the chunk has an
// odd number of Y, so Im q is needed as well.
template <typename P, int MS, int ADDR, bool DIAG, bool ODD>
__device__ __forceinline__ void etile_chunk(unsigned char const *tile_bytes, uint32_t lane_off, uint32_t x4,
                                            uint32_t n_blocks, uint32_t hp, uint32_t sh, uint32_t count,
                                            uint32_t const (&zk)[MS], uint32_t const (&zb)[MS],
                                            uint32_t const (&odd_ny)[MS], typename P::V (&r)[MS])
{
    uint32_t dk[8]; // ADDR == 2 only: byte offset of pair k inside the block
    if (ADDR == 2)
    {
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k)
            dk[k] = (((k >> hp) << (hp + 1)) | (k & ((1u << hp) - 1u))) << (sh + 4);
    }
    // blocks step the j bits above the in-block three; for ADDR == 2 the inserted zero lies below them
    uint32_t const hp_low = (1u << hp) - 1u;
    for (uint32_t blk = 0; blk < n_blocks; ++blk)
    {
        uint32_t const j0 = blk * 8;
        uint32_t const jp0 = ADDR == 2 ? (j0 << (sh + 1)) : ((((j0 >> hp) << (hp + 1)) | (j0 & hp_low)) << sh);
        uint32_t const a0 = lane_off + (jp0 << 4); // byte offset of the block's first row
        uint32_t const b0 = a0 ^ x4;
        using V = typename P::V;
        using Row = typename P::Row;
        V qre[8], qim[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            uint32_t const dko = ADDR == 0 ? (k << 9) : ADDR == 1 ? (k << 10) : dk[k];
            Row const a = *reinterpret_cast<Row const *>(tile_bytes + a0 + dko);
            V const ar = P::re(a), ai = P::im(a);
            if (DIAG)
                qre[k] = P::fma(ar, ar, P::mul(ai, ai));
            else
            {
                Row const b = *reinterpret_cast<Row const *>(tile_bytes + (b0 ^ dko));
                V const br = P::re(b), bi = P::im(b);
                qre[k] = P::fma(ar, br, P::mul(ai, bi));
                if (ODD)
                    qim[k] = P::fma(ar, bi, P::mul(ai, P::neg(br)));
            }
        }
#pragma unroll
        for (int m = 0; m < MS; ++m)
        {
            if (static_cast<uint32_t>(m) < count)
            {
                V rb;
                if (ODD && odd_ny[m])
                    rb = signed_sum8_dyn<P>(zk[m], qim);
                else
                    rb = signed_sum8_dyn<P>(zk[m], qre);
                r[m] = P::add(r[m], P::flip(rb, __popc(blk & zb[m]) & 1u));
            }
        }
    }
}

// Strings of one launch.  Whole-column launches (registers of <= 12 qubits): the masks are the global ones, one
// coset.  Coset launches (COSET = true, larger registers): the tile is one rank-12 coset of a pass of the coset plan
// (rows base ^ comb(basis, l)); x / z are the pass-local masks, the coset-base sign (-1)^{par(base & z)} is applied per
// string and coset, and the CTA walks all cosets for its column (pair), accumulating into E from the second coset on
// (the chunk -> warp assignment is fixed, so every E entry is updated by one thread in a fixed order).
struct EtStrings
{
    PairChunk const *chunks;
    uint32_t n_chunks;
    uint64_t const *sz;   // full z-mask per string
    uint32_t const *szl;  // pass-local z-mask (COSET only)
    uint8_t const *sodd;  // odd number of Y
    uint32_t const *sidx; // row of E (COSET only; identity otherwise)
    uint64_t basis[12];   // COSET only
    uint64_t nonpivot_mask;
    uint64_t n_cosets;
};

template <typename P, int MS, bool COSET>
__global__ void __launch_bounds__(kThreads)
    sop_expval_tile2_kernel(EtStrings st, uint32_t n_qubits, uint64_t rowvecs,
                            CVec<typename P::T, P::COLS> const *__restrict__ in, typename P::T *__restrict__ E /* [S][B] */,
                            uint64_t B)
{
    using V = typename P::V;
    using Row = typename P::Row;
    extern __shared__ __align__(16) unsigned char et_smem[];
    Row *tile = reinterpret_cast<Row *>(et_smem);
    uint32_t const tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t const rows = COSET ? 4096u : (1u << n_qubits);
    uint64_t const v = blockIdx.x; // column (pair) of this CTA: one 16-byte vector per row
    Row const *in4 = reinterpret_cast<Row const *>(in);
    PairChunk const *chunks = st.chunks;
    uint32_t const n_chunks = st.n_chunks;
    for (uint64_t coset = 0; coset < (COSET ? st.n_cosets : 1); ++coset)
    {
    uint64_t const base = COSET ? deposit_bits(coset, st.nonpivot_mask) : 0;
    if (COSET)
        __syncthreads(); // the previous coset's readers are done
    for (uint32_t r = tid; r < rows; r += kThreads)
    {
        uint64_t const grow = COSET ? (base ^ comb_of<12>(st.basis, r)) : r;
        tile[r] = P::to_row(in4[grow * rowvecs + v]); // complex64: planar per pair
    }
    __syncthreads();

    uint32_t const n_warps_total = (kThreads / 32) * gridDim.y;
    for (uint32_t c = blockIdx.y * (kThreads / 32) + warp; c < n_chunks; c += n_warps_total)
    {
        PairChunk const ch = chunks[c];
        uint32_t const x = static_cast<uint32_t>(ch.x);
        // pair index p = lane + 32 j  ->  row i = lane_part | jpart(j), where the pair's representative has bit hbit
        // of x cleared: the zero is inserted among the lane bits (hbit < 5) or among the j bits (hbit >= 5)
        uint32_t lane_part = lane, sh = 5, hp = 20; // hp: position of the inserted zero among the j bits (20: none)
        if (!ch.diag)
        {
            if (ch.hbit < 5)
            {
                lane_part = ((lane >> ch.hbit) << (ch.hbit + 1)) | (lane & ((1u << ch.hbit) - 1u));
                sh = 6;
            }
            else
                hp = ch.hbit - 5;
        }
        uint32_t const hp_low = (1u << hp) - 1u;
        uint32_t const n_blocks = (ch.diag ? rows : (rows >> 1)) >> 8; // 8 pairs per lane per block
        uint32_t sl[MS], zk[MS], zb[MS], odd_ny[MS];
        bool any_odd = false;
#pragma unroll
        for (int m = 0; m < MS; ++m)
        {
            bool const live = static_cast<uint32_t>(m) < ch.count;
            uint32_t const z = !live ? 0u : COSET ? st.szl[ch.s0 + m] : static_cast<uint32_t>(st.sz[ch.s0 + m]);
            odd_ny[m] = live ? st.sodd[ch.s0 + m] : 0u;
            any_odd |= odd_ny[m] != 0;
            sl[m] = __popc(lane_part & z) & 1u;
            if (COSET && live)
                sl[m] ^= parity64(base & st.sz[ch.s0 + m]);
            uint32_t const zz = z >> sh;
            uint32_t const zj = ((zz >> (hp + 1)) << hp) | (zz & hp_low); // z over the j bits
            zk[m] = zj & 7u;
            zb[m] = zj >> 3;
        }
        V r[MS];
#pragma unroll
        for (int m = 0; m < MS; ++m)
            r[m] = P::zero();

        unsigned char const *tb = et_smem;
        uint32_t const lane_off = lane_part << 4, x4 = x << 4;
        if (ch.diag)
            etile_chunk<P, MS, 0, true, false>(tb, lane_off, 0, n_blocks, hp, sh, ch.count, zk, zb, odd_ny, r);
        else if (ch.hbit < 5)
        {
            if (any_odd)
                etile_chunk<P, MS, 1, false, true>(tb, lane_off, x4, n_blocks, hp, sh, ch.count, zk, zb, odd_ny, r);
            else
                etile_chunk<P, MS, 1, false, false>(tb, lane_off, x4, n_blocks, hp, sh, ch.count, zk, zb, odd_ny, r);
        }
        else if (ch.hbit >= 8)
        {
            if (any_odd)
                etile_chunk<P, MS, 0, false, true>(tb, lane_off, x4, n_blocks, hp, sh, ch.count, zk, zb, odd_ny, r);
            else
                etile_chunk<P, MS, 0, false, false>(tb, lane_off, x4, n_blocks, hp, sh, ch.count, zk, zb, odd_ny, r);
        }
        else
        {
            if (any_odd)
                etile_chunk<P, MS, 2, false, true>(tb, lane_off, x4, n_blocks, hp, sh, ch.count, zk, zb, odd_ny, r);
            else
                etile_chunk<P, MS, 2, false, false>(tb, lane_off, x4, n_blocks, hp, sh, ch.count, zk, zb, odd_ny, r);
        }
#pragma unroll
        for (int m = 0; m < MS; ++m)
        {
            if (static_cast<uint32_t>(m) >= ch.count)
                break;
            V val = P::flip(r[m], sl[m]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
                val = P::shfl_add(val, off);
            if (lane == 0)
                P::store(E, COSET ? st.sidx[ch.s0 + m] : ch.s0 + m, B, v, val, COSET && coset > 0);
        }
    }
    }
}

} // namespace fpk
