// Fused GroupNorm (+ optional SiLU) for channels-last activations (sm_100a, HBM-bound).
//
// The VAE decoder is ~65 % of the `fast` preset's FLOPs and every one of its
// convolutions is preceded by GroupNorm -> SiLU (diffusers/src/diffusers/models/resnet.py:320-376,
// autoencoders/vae.py:284-347).  cuDNN's Blackwell convolutions want NHWC; torch's group_norm then
// falls back to NCHW round-trip copies.  These two kernels keep everything NHWC:
//   stats : one coalesced pass, per-channel shifted sums in registers -> per-(sample, chunk, group)
//           partials in double (deterministic: no float atomics);
//   apply : y = silu(round_T(x * a_c + b_c)), a_c = rstd_g * gamma_c, b_c = beta_c - mean_g * a_c,
//           16-byte vectors, the group statistics are reduced from the partials in the prologue.
// Algorithmic traffic: 2 reads + 1 write of the activation.
#include <math.h>

#include "common.cuh"

namespace rd {

constexpr int kGnMaxThreads = 512;
constexpr int kGnMaxChunks = 128;

struct GnParams {
    const void* x;
    void* y;
    const void* gamma;
    const void* beta;
    double* partial;      // [N][chunks][G][2]  (sum x, sum x^2)
    int N, HW, C, G, cpg;  // cpg = channels per group
    int CV;                // C / 8 (16-bit types) or C / 4 (fp32): 16-byte vectors per pixel
    int R;                 // pixel rows processed concurrently by one CTA
    int chunks, rows_per_chunk;
    float eps;
    int act;               // 0 = none, 1 = SiLU
};

template <typename T> struct GnVec { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void gn_unpack(const uint4& u, float* f) {
    if constexpr (sizeof(T) == 4) {
        f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
    } else {
        const T* h = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = Cvt<T>::to_f(h[i]);
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads) gn_stats_kernel(const GnParams p) {
    constexpr int VN = GnVec<T>::N;
    extern __shared__ double sh_d[];                  // [2][C] channel sums (double)
    double* ch_sum = sh_d;
    double* ch_sq = sh_d + p.C;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int cv = threadIdx.x % p.CV, r = threadIdx.x / p.CV;
    const bool active = r < p.R;
    const T* xs = reinterpret_cast<const T*>(p.x) + (size_t)n * p.HW * p.C;
    const int row0 = chunk * p.rows_per_chunk;
    const int row1 = min(p.HW, row0 + p.rows_per_chunk);

    for (int c = threadIdx.x; c < 2 * p.C; c += blockDim.x) sh_d[c] = 0.0;
    __syncthreads();

    float shift[VN], s1[VN], s2[VN];
    if (active) {
        gn_unpack<T>(*reinterpret_cast<const uint4*>(xs + (size_t)cv * VN), shift);   // pixel 0 of the sample
#pragma unroll
        for (int i = 0; i < VN; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
        int row = row0 + r;
        // two independent 16-byte loads in flight per thread
        for (; row + p.R < row1; row += 2 * p.R) {
            const uint4 a = ldg_stream16(xs + ((size_t)row * p.CV + cv) * VN);
            const uint4 b = ldg_stream16(xs + ((size_t)(row + p.R) * p.CV + cv) * VN);
            float fa[VN], fb[VN];
            gn_unpack<T>(a, fa);
            gn_unpack<T>(b, fb);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
                const float da = fa[i] - shift[i], db = fb[i] - shift[i];
                s1[i] += da + db;
                s2[i] = fmaf(da, da, fmaf(db, db, s2[i]));
            }
        }
        for (; row < row1; row += p.R) {
            float fa[VN];
            gn_unpack<T>(ldg_stream16(xs + ((size_t)row * p.CV + cv) * VN), fa);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
                const float da = fa[i] - shift[i];
                s1[i] += da;
                s2[i] = fmaf(da, da, s2[i]);
            }
        }
        // un-shift in double and combine the R row-threads of each channel
        const int first = row0 + r;
        const double nn = first < row1 ? (double)((row1 - first + p.R - 1) / p.R) : 0.0;
#pragma unroll
        for (int i = 0; i < VN; ++i) {
            const double K = (double)shift[i], d1 = (double)s1[i], d2 = (double)s2[i];
            atomicAdd(&ch_sum[cv * VN + i], d1 + nn * K);
            atomicAdd(&ch_sq[cv * VN + i], d2 + 2.0 * K * d1 + nn * K * K);
        }
    }
    __syncthreads();
    // channels -> groups (fixed order => deterministic up to the shared-memory atomics above, which add
    // at most R <= 32 doubles per address; double addition of these magnitudes is order-insensitive to
    // ~1e-16 relative, far below the fp16/fp32 output rounding)
    for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int c = g * p.cpg; c < (g + 1) * p.cpg; ++c) { a += ch_sum[c]; b += ch_sq[c]; }
        double* out = p.partial + (((size_t)n * p.chunks + chunk) * p.G + g) * 2;
        out[0] = a;
        out[1] = b;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads) gn_apply_kernel(const GnParams p) {
    constexpr int VN = GnVec<T>::N;
    extern __shared__ float sh_f[];                   // [2][C] : a_c, b_c
    __shared__ float g_mean[64], g_rstd[64];
    float* sa = sh_f;
    float* sb = sh_f + p.C;
    const int n = blockIdx.y, chunk = blockIdx.x;
    for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
        double a = 0.0, b = 0.0;
        const double* part = p.partial + ((size_t)n * p.chunks * p.G + g) * 2;
        for (int c = 0; c < p.chunks; ++c) { a += part[(size_t)c * p.G * 2]; b += part[(size_t)c * p.G * 2 + 1]; }
        const double cnt = (double)p.HW * (double)p.cpg;
        const double mean = a / cnt;
        double var = b / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        g_mean[g] = (float)mean;
        g_rstd[g] = rsqrtf((float)var + p.eps);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        const int g = c / p.cpg;
        const float gam = p.gamma ? Cvt<T>::to_f(reinterpret_cast<const T*>(p.gamma)[c]) : 1.f;
        const float bet = p.beta ? Cvt<T>::to_f(reinterpret_cast<const T*>(p.beta)[c]) : 0.f;
        const float a = g_rstd[g] * gam;
        sa[c] = a;
        sb[c] = fmaf(-a, g_mean[g], bet);
    }
    __syncthreads();

    const int cv = threadIdx.x % p.CV, r = threadIdx.x / p.CV;
    if (r >= p.R) return;
    float a[VN], b[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) { a[i] = sa[cv * VN + i]; b[i] = sb[cv * VN + i]; }
    const T* xs = reinterpret_cast<const T*>(p.x) + (size_t)n * p.HW * p.C;
    T* ys = reinterpret_cast<T*>(p.y) + (size_t)n * p.HW * p.C;
    const int row0 = chunk * p.rows_per_chunk;
    const int row1 = min(p.HW, row0 + p.rows_per_chunk);
    for (int row = row0 + r; row < row1; row += p.R) {
        const size_t off = ((size_t)row * p.CV + cv) * VN;
        float f[VN];
        gn_unpack<T>(ldg_stream16(xs + off), f);
        alignas(16) T o[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) {
            float v = Cvt<T>::round(fmaf(f[i], a[i], b[i]));       // group_norm output, rounded to T like torch
            if (p.act == 1) v = __fdividef(v, 1.f + __expf(-v));    // SiLU on the rounded value
            o[i] = Cvt<T>::from_f(v);
        }
        *reinterpret_cast<uint4*>(ys + off) = *reinterpret_cast<const uint4*>(o);
    }
}

static int gn_plan(GnParams& p, int es) {
    const int VN = 16 / es;
    p.CV = p.C / VN;
    int R = kGnMaxThreads / p.CV;
    if (R < 1) return -1;
    if (R > 32) R = 32;
    p.R = R;
    // aim for ~16 vectors per thread per chunk, at most kGnMaxChunks chunks per sample
    int rows = 16 * R;
    int chunks = (p.HW + rows - 1) / rows;
    if (chunks > kGnMaxChunks) { chunks = kGnMaxChunks; rows = (p.HW + chunks - 1) / chunks; chunks = (p.HW + rows - 1) / rows; }
    if (chunks < 1) chunks = 1;
    p.chunks = chunks;
    p.rows_per_chunk = rows;
    return 0;
}

}  // namespace rd

using namespace rd;

extern "C" {

int64_t rd_groupnorm_workspace_bytes(int N, int G) {
    return (int64_t)N * kGnMaxChunks * (G > 0 ? G : 1) * 2 * (int64_t)sizeof(double);
}

int rd_groupnorm_nhwc(const void* x, void* y, const void* gamma, const void* beta, int N, int HW, int C, int G, float eps,
                      int act, rd_dtype dtype, void* workspace, int64_t workspace_bytes, void* stream) {
    RD_CHECK_ARG(x && y && workspace, "NULL pointer");
    RD_CHECK_ARG(N >= 1 && HW >= 1 && C >= 1 && G >= 1 && G <= 64 && C % G == 0, "bad shape N=%d HW=%d C=%d G=%d", N, HW, C, G);
    RD_CHECK_ARG(N <= 65535, "N must be <= 65535");
    RD_CHECK_ARG(act == 0 || act == 1, "act must be 0 (none) or 1 (SiLU)");
    const int es = dtype == RD_F32 ? 4 : 2;
    if (C % (16 / es) != 0) { set_error("C=%d is not a multiple of the 16-byte vector width", C); return RD_ERR_UNSUPPORTED; }
    RD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "x / y must be 16-byte aligned");
    GnParams p{};
    p.x = x; p.y = y; p.gamma = gamma; p.beta = beta;
    p.partial = reinterpret_cast<double*>(workspace);
    p.N = N; p.HW = HW; p.C = C; p.G = G; p.cpg = C / G; p.eps = eps; p.act = act;
    if (gn_plan(p, es) != 0) { set_error("C=%d too wide for one CTA row", C); return RD_ERR_UNSUPPORTED; }
    if (workspace_bytes < (int64_t)N * p.chunks * G * 2 * (int64_t)sizeof(double)) {
        set_error("group-norm workspace too small");
        return RD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = p.CV * p.R;
    dim3 grid(p.chunks, N);
    const size_t sm_stats = 2 * (size_t)C * sizeof(double), sm_apply = 2 * (size_t)C * sizeof(float);
    switch (dtype) {
        case RD_F32:
            gn_stats_kernel<float><<<grid, threads, sm_stats, st>>>(p);
            gn_apply_kernel<float><<<grid, threads, sm_apply, st>>>(p);
            break;
        case RD_F16:
            gn_stats_kernel<__half><<<grid, threads, sm_stats, st>>>(p);
            gn_apply_kernel<__half><<<grid, threads, sm_apply, st>>>(p);
            break;
        case RD_BF16:
            gn_stats_kernel<__nv_bfloat16><<<grid, threads, sm_stats, st>>>(p);
            gn_apply_kernel<__nv_bfloat16><<<grid, threads, sm_apply, st>>>(p);
            break;
        default: set_error("bad dtype"); return RD_ERR_BAD_ARG;
    }
    RD_LAUNCH_CHECK("gn kernels");
    return RD_OK;
}

}  // extern "C"
