// Standalone A/B harness for the coset apply kernels (compiles in seconds, unlike the whole library):
// builds a random operator, plans it with the product planner, runs coset_kernel (K3b) and coset_few_kernel (K3e)
// on the same device-resident batch, compares the results and times both with CUDA events.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
//             -I fast-pauli_b200/csrc -o scripts/micro/coset_bench scripts/micro/coset_bench.cu
// run:   coset_bench <case: few|rand|few16> <n_qubits> <B> <ctPerCta> [log_twc] [scatter 0|1]
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "coset.cuh"
#include "coset2.cuh"
#include "coset3.cuh"
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
    int scatter = argc > 6 ? atoi(argv[6]) : 0;
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
    else if (kind.rfind("rand", 0) == 0 && kind.size() > 4)
        variants(atoi(kind.c_str() + 4), 1); // randN: N random strings
    else if (kind == "cfg3")
    {
        // BASELINE config 3's operator: 2000 strings of weight <= 4 (weight uniform in 1..4, positions without replacement)
        for (int t = 0; t < 2000; ++t)
        {
            std::vector<uint8_t> s(n, 0);
            int w = 1 + static_cast<int>(rng() % 4);
            for (int k = 0; k < w; ++k)
            {
                int pos;
                do
                    pos = static_cast<int>(rng() % n);
                while (s[pos]);
                s[pos] = 1 + static_cast<uint8_t>(rng() % 3);
            }
            push(s);
            ++S;
        }
    }
    else if (kind == "chain")
    {
        for (int i = 0; i + 1 < n; ++i)
            for (uint8_t pp : {1, 2, 3})
            {
                std::vector<uint8_t> s(n, 0);
                s[i] = pp;
                s[i + 1] = pp;
                push(s);
                ++S;
            }
    }
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
    int const nbuf = scatter; // 7th argument: 1 = single buffer, 2 = double buffer; 8th: min CTAs/SM; 9th: TMA fill
    int const minb = argc > 7 ? atoi(argv[7]) : 2;
    int const tma = argc > 8 ? atoi(argv[8]) : 0;
    (void)minb;
    std::vector<FewStrings<T>> pstr(views.size());
    bool pstr_ok = true;
    for (size_t p = 0; p < passes.size(); ++p)
    {
        auto const &hp = passes[p];
        memset(&pstr[p], 0, sizeof(FewStrings<T>));
        if (hp.gxl.size() > 8 || hp.sz.size() > kFewParamStrings)
        {
            pstr_ok = false;
            continue;
        }
        for (size_t i = 0; i < hp.sz.size(); ++i)
        {
            pstr[p].c[i] = Cx<T>{hp.sc[i].real(), hp.sc[i].imag()};
            pstr[p].z[i] = hp.sz[i];
        }
        for (size_t g = 0; g <= hp.gxl.size(); ++g)
            pstr[p].gs[g] = hp.gstart[g];
        for (size_t g = 0; g < hp.gxl.size(); ++g)
            pstr[p].gxl[g] = hp.gxl[g];
    }
    bool const use_pstr = tma && pstr_ok;
    (void)use_pstr; // 9th argument now selects the parameter-block row-factor phase
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, cuuint64_t const *, cuuint64_t const *,
                                 cuuint32_t const *, cuuint32_t const *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void **>(&enc), cudaEnableDefault, &qres));
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    {
        cuuint64_t dims[2] = {(cuuint64_t)B * 2, (cuuint64_t)dim};
        cuuint64_t strides[1] = {(cuuint64_t)B * 16};
        cuuint32_t box[2] = {32, 1};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
        {
            printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
            exit(3);
        }
    }
    int const grid_f = minb > 2 ? minb : 148; // K3f: 8th argument > 2 overrides the persistent grid size
    auto run_new = [&](double *out) {
        for (size_t p = 0; p < views.size(); ++p)
        {
            if (nbuf == 5)
            {
                // K3i: direct-store TMA-fed kernel (one string per x-mask)
                auto const &hp = passes[p];
                if (hp.sz.size() != hp.gxl.size() || hp.gxl.size() > (size_t)kDirMaxMasks)
                {
                    printf("pass not eligible for K3i\n");
                    exit(1);
                }
                DirStrings<T> ds;
                memset(&ds, 0, sizeof ds);
                ds.n = (uint32_t)hp.gxl.size();
                for (size_t g = 0; g < hp.gxl.size(); ++g)
                {
                    ds.c[g] = Cx<T>{hp.sc[g].real(), hp.sc[g].imag()};
                    ds.z[g] = hp.sz[g];
                    ds.xl[g] = hp.gxl[g];
                    ds.zl[g] = hp.szl[g];
                }
                size_t const smem = kFewTmaBufs * kFewTmaTile;
                uint32_t const nct = (uint32_t)(rowvecs >> 4);
                uint64_t const n_tiles = (dim >> 8) * nct;
                unsigned const g = (unsigned)std::min<uint64_t>(grid_f, n_tiles);
                int const nch = (int)((ds.n + 7) / 8);
#define LAUNCH_DIR2(NCH, IB)                                                                                           \
    {                                                                                                                  \
        CK(cudaFuncSetAttribute(coset_dir_tma_kernel<T, 1, NCH, IB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        coset_dir_tma_kernel<T, 1, NCH, IB><<<g, kDirThreads, smem>>>(views[p], rowvecs, nct, n_tiles,                 \
                                                                      reinterpret_cast<Vec *>(out), p ? 1 : 0, ds, tm); \
    }
#define LAUNCH_DIR(NCH)                                                                                                \
    {                                                                                                                  \
        if (log_twc == 8)                                                                                              \
            LAUNCH_DIR2(NCH, 8)                                                                                        \
        else if (log_twc == 2)                                                                                         \
            LAUNCH_DIR2(NCH, 2)                                                                                        \
        else                                                                                                           \
            LAUNCH_DIR2(NCH, 4)                                                                                        \
    }
                if (nch <= 1)
                    LAUNCH_DIR(1)
                else if (nch == 2)
                    LAUNCH_DIR(2)
                else
                    LAUNCH_DIR(4)
                continue;
            }
            if (nbuf == 4)
            {
                size_t const smem = kFewTmaBufs * kFewTmaTile + kGenMetaBytes;
                CK(cudaFuncSetAttribute(coset_gen_tma_kernel<T, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                uint64_t const n_pairs = (dim >> 8) / 2;
                uint32_t const nct = (uint32_t)(rowvecs >> 4);
                uint32_t chunk = ctPer > 0 ? (uint32_t)ctPer : nct;
                while (nct % chunk)
                    --chunk;
                uint64_t const items = n_pairs * (nct / chunk);
                unsigned const g = (unsigned)std::min<uint64_t>(grid_f, items);
                static std::vector<GenStrings<T>> gstrs;
                if (gstrs.empty())
                {
                    gstrs.resize(passes.size());
                    for (size_t q = 0; q < passes.size(); ++q)
                    {
                        auto const &hp = passes[q];
                        memset(&gstrs[q], 0, sizeof(GenStrings<T>));
                        if (hp.sz.size() > kGenMaxStrings || hp.gxl.size() > kGenMaxGroups)
                        {
                            printf("pass too large for K3g\n");
                            exit(1);
                        }
                        for (size_t i = 0; i < hp.sz.size(); ++i)
                        {
                            gstrs[q].c[i] = Cx<T>{hp.sc[i].real(), hp.sc[i].imag()};
                            gstrs[q].z[i] = (uint32_t)hp.sz[i];
                        }
                        for (size_t gg = 0; gg <= hp.gxl.size(); ++gg)
                            gstrs[q].gs[gg] = (uint16_t)hp.gstart[gg];
                        for (size_t gg = 0; gg < hp.gxl.size(); ++gg)
                            gstrs[q].gxl[gg] = (uint8_t)hp.gxl[gg];
                    }
                }
                if (tma)
                {
                    CK(cudaFuncSetAttribute(coset_gen_tma_kernel<T, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    coset_gen_tma_kernel<T, 1, true><<<g, kFewTmaThreads, smem>>>(views[p], rowvecs, nct, n_pairs, chunk,
                                                                                  reinterpret_cast<Vec *>(out), p ? 1 : 0, tm, gstrs[p]);
                }
                else
                    coset_gen_tma_kernel<T, 1, false><<<g, kFewTmaThreads, smem>>>(views[p], rowvecs, nct, n_pairs, chunk,
                                                                                   reinterpret_cast<Vec *>(out), p ? 1 : 0, tm, gstrs[p]);
                continue;
            }
            if (nbuf == 3)
            {
                size_t const smem = kFewTmaBufs * kFewTmaTile;
                CK(cudaFuncSetAttribute(coset_few_tma_kernel<T, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                uint64_t const n_pairs = (dim >> 8) / 2;
                unsigned const g = (unsigned)std::min<uint64_t>(grid_f, n_pairs);
                coset_few_tma_kernel<T, 1, 8><<<g, kFewTmaThreads, smem>>>(views[p], rowvecs, (uint32_t)(rowvecs >> 4), n_pairs,
                                                                           reinterpret_cast<Vec *>(out), p ? 1 : 0, pstr[p], tm);
                continue;
            }
#define LAUNCH_FEW(LT, NB, MB, TM)                                                                                     \
    if (ctPer < 0)                                                                                                     \
    {                                                                                                                  \
        using Cfg = FewCfg<LT>;                                                                                        \
        size_t const sm = (NB + 1) * Cfg::TILE_BYTES;                                                                  \
        CK(cudaFuncSetAttribute(coset_few_kernel<T, 1, LT, 8, NB, MB, TM, true>,                                       \
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));                                \
        uint32_t const nct = rowvecs >> LT;                                                                            \
        coset_few_kernel<T, 1, LT, 8, NB, MB, TM, true><<<(unsigned)(dim >> 8), 256, sm>>>(                            \
            views[p], rowvecs, nct, nct, 1, reinterpret_cast<Vec const *>(in), reinterpret_cast<Vec *>(out), p ? 1 : 0, \
            pstr[p]);                                                                                                  \
    }                                                                                                                  \
    else                                                                                                               \
    {                                                                                                                  \
        using Cfg = FewCfg<LT>;                                                                                        \
        CK(cudaFuncSetAttribute(coset_few_kernel<T, 1, LT, 8, NB, MB, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)(NB * Cfg::TILE_BYTES)));                                                         \
        uint32_t const nct = rowvecs >> LT;                                                                            \
        uint32_t per = ctPer > 0 ? std::min<uint32_t>(ctPer, nct) : nct;                                               \
        uint32_t groups = (nct + per - 1) / per;                                                                       \
        coset_few_kernel<T, 1, LT, 8, NB, MB, TM><<<(unsigned)((dim >> 8) * groups), 256, NB * Cfg::TILE_BYTES>>>(     \
            views[p], rowvecs, nct, per, groups, reinterpret_cast<Vec const *>(in), reinterpret_cast<Vec *>(out),      \
            p ? 1 : 0, pstr[p]);                                                                                            \
    }
            bool const tma = use_pstr;
            if (log_twc == 4 && nbuf <= 1 && !tma)
                LAUNCH_FEW(4, 1, 2, false)
            else if (log_twc == 4 && nbuf <= 1 && tma)
                LAUNCH_FEW(4, 1, 2, true)
            else if (log_twc == 3 && nbuf <= 1 && !tma)
                LAUNCH_FEW(3, 1, 2, false)
            else if (log_twc == 3 && nbuf == 2 && !tma)
                LAUNCH_FEW(3, 2, 2, false)
            else if (log_twc == 3 && nbuf == 2 && tma)
                LAUNCH_FEW(3, 2, 2, true)
            else if (log_twc == 3 && nbuf <= 1 && tma)
                LAUNCH_FEW(3, 1, 2, true)
            else
            {
                printf("unsupported variant\n");
                exit(1);
            }
        }
    };
    bool const few_ok = [&]() {
        for (auto const &v : views)
            if (v.n_groups > 8 && nbuf != 4 && nbuf != 5)
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
    {
#ifdef FP_FEW_PROFILE
        unsigned long long z[8] = {};
        CK(cudaMemcpyToSymbol(g_few_prof, z, sizeof z));
#endif
        time_it(run_new, out_b, "K3e");
#ifdef FP_DIR_PROFILE
        {
            unsigned long long z[8] = {};
            CK(cudaMemcpyFromSymbol(z, g_dir_prof, sizeof z));
            double const tiles = 8.0 * views.size() * (dim >> 8) * (rowvecs >> 4); // 1 check + 2 warm-up + 5 timed runs
            printf("  K3i per tile and warp (cycles): setup %.0f  wait-full %.0f  compute+store %.0f  fence+arrive %.0f | producer: wait-empty %.0f issue %.0f\n",
                   z[0] / tiles / 16, z[1] / tiles / 16, z[2] / tiles / 16, z[3] / tiles / 16, z[4] / tiles, z[5] / tiles);
        }
#endif
#ifdef FP_FEW_PROFILE
        CK(cudaMemcpyFromSymbol(z, g_few_prof, sizeof z));
        double const ctas = 7.0 * views.size() * (dim >> 8) * ((rowvecs >> log_twc) / std::max(1, ctPer ? ctPer : (int)(rowvecs >> log_twc)));
        double const tiles = 7.0 * views.size() * (dim >> 8) * (rowvecs >> log_twc);
        printf("  phases (cycles): setup+D %.0f /CTA | per tile: wait-fill %.0f  compute %.0f  stage %.0f  store %.0f  refill-issue %.0f\n",
               z[0] / ctas, z[1] / tiles, z[2] / tiles, z[3] / tiles, z[4] / tiles, z[5] / tiles);
#endif
    }
    return 0;
}
