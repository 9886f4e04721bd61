// Standalone A/B harness for the coset apply kernels (compiles in seconds, unlike the whole library):
// builds a random operator, plans it with the product planner, runs coset_kernel (K3b) and coset_few_kernel (K3e)
// on the same device-resident batch, compares the results and times both with CUDA events.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
//             -I fast-pauli_b200/csrc -o scripts/micro/coset_bench scripts/micro/coset_bench.cu
// run:   coset_bench <case: few|rand|few16> <n_qubits> <B> <ctPerCta> [log_twc]
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "coset.cuh"
#include "coset2.cuh"
#include "coset_plan.hpp"

using namespace fpk;

#define CK(x)                                                                                                          \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess)                                                                                         \
        {                                                                                                              \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                            \
            exit(2);                                                                                                   \
        }                                                                                                              \
    } while (0)

template <typename V> V *upload(std::vector<V> const &v)
{
    V *p;
    CK(cudaMalloc(&p, std::max<size_t>(16, v.size() * sizeof(V))));
    CK(cudaMemcpy(p, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
    return p;
}

__global__ void k_fill(double *p, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = static_cast<double>((i * 2654435761ull) & 0xffffff) * (1.0 / 16777216.0) - 0.5;
}
__global__ void k_maxdiff(double const *a, double const *b, size_t n, double *res)
{
    double m = 0, s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        m = fmax(m, fabs(a[i] - b[i]));
        s = fmax(s, fabs(a[i]));
    }
    atomicMax(reinterpret_cast<unsigned long long *>(res), __double_as_longlong(m));
    atomicMax(reinterpret_cast<unsigned long long *>(res + 1), __double_as_longlong(s));
}

int main(int argc, char **argv)
{
    std::string kind = argc > 1 ? argv[1] : "few";
    int n = argc > 2 ? atoi(argv[2]) : 20;
    int B = argc > 3 ? atoi(argv[3]) : 64;
    int ctPer = argc > 4 ? atoi(argv[4]) : 0;
    int log_twc = argc > 5 ? atoi(argv[5]) : 4;
    using T = double;

    std::mt19937_64 rng(1234);
    std::vector<uint8_t> codes;
    auto push = [&](std::vector<uint8_t> const &s) { codes.insert(codes.end(), s.begin(), s.end()); };
    size_t S = 0;
    auto rand_string = [&]() {
        std::vector<uint8_t> s(n);
        for (auto &c : s)
            c = rng() & 3;
        return s;
    };
    auto variants = [&](int masks, int per) {
        for (int m = 0; m < masks; ++m)
        {
            auto s = rand_string();
            for (int v = 0; v < per; ++v)
            {
                auto t = s;
                for (auto &c : t)
                    if (rng() & 1)
                        c = (c == 1) ? 2 : (c == 2) ? 1 : (c == 0) ? 3 : 0; // X<->Y, I<->Z keeps the x-mask
                push(t);
                ++S;
            }
        }
    };
    if (kind == "few")
        variants(8, 8);
    else if (kind == "few16")
        variants(16, 4);
    else if (kind == "few4")
        variants(4, 16);
    else
        variants(64, 1);
    std::vector<std::complex<T>> h(S);
    for (auto &c : h)
        c = {std::uniform_real_distribution<double>(-1, 1)(rng), std::uniform_real_distribution<double>(-1, 1)(rng)};
    PackedOp<T> op = pack_op<T>(n, S, codes.data(), h.data(), true);
    std::vector<CosetPassHost<T>> passes = plan_coset<T>(op, n, 8, 0);
    printf("%s: n=%d B=%d strings=%zu groups=%zu passes=%zu\n", kind.c_str(), n, B, S, op.gx.size(), passes.size());

    std::vector<CosetPassView<T>> views;
    for (auto const &hp : passes)
    {
        CosetPassView<T> v{};
        for (int k = 0; k < kCosetMaxRank; ++k)
            v.basis[k] = k < hp.basis.r ? hp.basis.b[k] : 0;
        v.nonpivot_mask = hp.nonpivot_mask;
        v.chunks = upload(hp.chunks);
        v.gxl = upload(hp.gxl);
        v.gstart = upload(hp.gstart);
        v.szl = upload(hp.szl);
        v.sz = upload(hp.sz);
        std::vector<Cx<T>> sc(hp.sc.size());
        for (size_t i = 0; i < sc.size(); ++i)
            sc[i] = Cx<T>{hp.sc[i].real(), hp.sc[i].imag()};
        v.scoef = upload(sc);
        v.sidx = upload(hp.sidx);
        v.n_chunks = static_cast<uint32_t>(hp.chunks.size());
        v.n_groups = static_cast<uint32_t>(hp.gxl.size());
        views.push_back(v);
        printf("  pass: groups=%u strings=%zu\n", v.n_groups, hp.sz.size());
    }

    size_t const dim = size_t(1) << n, n_dbl = dim * B * 2;
    double *in, *out_a, *out_b, *res;
    CK(cudaMalloc(&in, n_dbl * 8));
    CK(cudaMalloc(&out_a, n_dbl * 8));
    CK(cudaMalloc(&out_b, n_dbl * 8));
    CK(cudaMalloc(&res, 16));
    CK(cudaMemset(res, 0, 16));
    k_fill<<<1184, 256>>>(in, n_dbl);
    uint64_t const rowvecs = B;
    using Vec = CVec<T, 1>;

    auto run_old = [&](double *out) {
        constexpr int LT = 4;
        size_t const smem = coset_smem_bytes<T, LT, 8, 16>();
        CK(cudaFuncSetAttribute(coset_kernel<T, 1, LT, 8, 0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t const nct = rowvecs >> LT;
        for (size_t p = 0; p < views.size(); ++p)
            coset_kernel<T, 1, LT, 8, 0, 16><<<(unsigned)((dim >> 8) * nct), 256, smem>>>(
                views[p], rowvecs, nct, reinterpret_cast<Vec const *>(in), reinterpret_cast<Vec *>(out), p ? 1 : 0, nullptr, 0,
                nullptr, nullptr, B);
    };
    auto run_new = [&](double *out) {
        for (size_t p = 0; p < views.size(); ++p)
        {
            if (log_twc == 4)
            {
                using Cfg = FewCfg<4>;
                CK(cudaFuncSetAttribute(coset_few_kernel<T, 1, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::TILE_BYTES));
                uint32_t const nct = rowvecs >> 4;
                uint32_t per = ctPer > 0 ? std::min<uint32_t>(ctPer, nct) : nct;
                uint32_t groups = (nct + per - 1) / per;
                coset_few_kernel<T, 1, 4, 8><<<(unsigned)((dim >> 8) * groups), 256, Cfg::TILE_BYTES>>>(
                    views[p], rowvecs, nct, per, groups, reinterpret_cast<Vec const *>(in), reinterpret_cast<Vec *>(out), p ? 1 : 0);
            }
            else
            {
                using Cfg = FewCfg<3>;
                CK(cudaFuncSetAttribute(coset_few_kernel<T, 1, 3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::TILE_BYTES));
                uint32_t const nct = rowvecs >> 3;
                uint32_t per = ctPer > 0 ? std::min<uint32_t>(ctPer, nct) : nct;
                uint32_t groups = (nct + per - 1) / per;
                coset_few_kernel<T, 1, 3, 8><<<(unsigned)((dim >> 8) * groups), 256, Cfg::TILE_BYTES>>>(
                    views[p], rowvecs, nct, per, groups, reinterpret_cast<Vec const *>(in), reinterpret_cast<Vec *>(out), p ? 1 : 0);
            }
        }
    };
    bool const few_ok = [&]() {
        for (auto const &v : views)
            if (v.n_groups > 8)
                return false;
        return true;
    }();
    run_old(out_a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (few_ok)
    {
        run_new(out_b);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        k_maxdiff<<<1184, 256>>>(out_a, out_b, n_dbl, res);
        double hres[2];
        CK(cudaMemcpy(hres, res, 16, cudaMemcpyDeviceToHost));
        printf("  max |old - new| = %.3e (max |old| = %.3e)\n", hres[0], hres[1]);
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time_it = [&](auto &&f, double *out, char const *name) {
        f(out);
        f(out);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 5; ++i)
            f(out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 5;
        printf("  %-8s %.3f ms  %.0f GB/s algorithmic (32 B/amp)\n", name, ms, dim * B * 32.0 / (ms * 1e-3) / 1e9);
    };
    time_it(run_old, out_a, "K3b");
    if (few_ok)
        time_it(run_new, out_b, "K3e");
    return 0;
}
