// Fused GroupNorm (+ optional SiLU) for channels-last activations (sm_100a, HBM-bound).
//
// The VAE decoder is ~65 % of the `fast` preset's FLOPs and every one of its
// convolutions is preceded by GroupNorm -> SiLU (diffusers/src/diffusers/models/resnet.py:320-376,
// autoencoders/vae.py:284-347).  cuDNN's Blackwell convolutions want NHWC; torch's group_norm then
// falls back to NCHW round-trip copies.  These two kernels keep everything NHWC:
//   stats : one coalesced pass, per-channel shifted sums in registers -> per-(sample, chunk, group)
//           partials in double.  No atomics of any kind: the row-threads of a channel are combined through
//           shared memory in a fixed order, so two launches on the same input give the same bits;
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
    const void* pre;      // optional [N][C] vector added to x before the statistics (conv bias + time embedding)
    double* partial;      // [N][chunks][G][2]  (sum x, sum x^2)
    int N, HW, C, G, cpg;  // cpg = channels per group
    int Cs;                // channels per CTA slice (blockIdx.z selects the slice); Cs % cpg == 0
    int CV;                // Cs / 8 (16-bit types) or Cs / 4 (fp32): 16-byte vectors per pixel and slice
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

// ---- packed fp32x2 element math (sm_100: FFMA2 / FADD2 / FMUL2 do two lanes in one issue slot).  These kernels are not only
// bandwidth- but also energy-limited inside a step (the step runs at the board's power cap): the pair forms below do the same
// fp32 arithmetic with the same roundings as the scalar code they replace in ~7 instead of ~12.5 issue slots per element
// (no range fix-up around ex2: 1 + ex2(t) is 1 for every t the fix-up would have touched; v * rcp(d) instead of div.approx).
__device__ __forceinline__ uint64_t gn_pk(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 gn_upk(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ uint64_t gn_ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t gn_fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t gn_fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float gn_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gn_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// two adjacent elements of T <-> one packed fp32 pair (one conversion instruction per pair when packing)
template <typename T> struct GnPair;
template <> struct GnPair<__half> {
    static __device__ __forceinline__ uint64_t unpack(uint32_t w) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        return gn_pk(f.x, f.y);
    }
    static __device__ __forceinline__ uint32_t pack(uint64_t v) {
        const float2 f = gn_upk(v);
        const __half2 h = __floats2half2_rn(f.x, f.y);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
};
template <> struct GnPair<__nv_bfloat16> {
    static __device__ __forceinline__ uint64_t unpack(uint32_t w) {
        return gn_pk(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    }
    static __device__ __forceinline__ uint32_t pack(uint64_t v) {
        const float2 f = gn_upk(v);
        const __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
};
// SiLU of a pair: v / (1 + exp(-v)) as v * rcp(1 + ex2(-v * log2 e))
__device__ __forceinline__ uint64_t gn_silu2(uint64_t v) {
    const float2 t = gn_upk(gn_fmul2(v, gn_pk(-1.4426950408889634f, -1.4426950408889634f)));
    const float2 d = gn_upk(gn_fadd2(gn_pk(gn_ex2(t.x), gn_ex2(t.y)), gn_pk(1.f, 1.f)));
    return gn_fmul2(v, gn_pk(gn_rcp(d.x), gn_rcp(d.y)));
}

// One 16-byte vector of the normalise(+SiLU) pass: y = act(round_T(round_T(x + pre) * a + b)); a2 / b2 / pre2 hold the per-channel
// constants of the vector's VN channels as VN / 2 packed pairs.
template <typename T, bool HAS_PRE>
__device__ __forceinline__ uint4 gn_apply_vec(const uint4& in, const uint64_t* a2, const uint64_t* b2, const uint64_t* pre2, int act) {
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};
    uint32_t o[4];
    if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint64_t x = GnPair<T>::unpack(w[j]);
            if (HAS_PRE) x = GnPair<T>::unpack(GnPair<T>::pack(gn_fadd2(x, pre2[j])));
            uint32_t h = GnPair<T>::pack(gn_ffma2(x, a2[j], b2[j]));          // group_norm output, rounded to T like torch
            if (act == 1) h = GnPair<T>::pack(gn_silu2(GnPair<T>::unpack(h)));  // SiLU on the rounded value
            o[j] = h;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint64_t x = gn_pk(__uint_as_float(w[2 * j]), __uint_as_float(w[2 * j + 1]));
            if (HAS_PRE) x = gn_fadd2(x, pre2[j]);
            uint64_t y = gn_ffma2(x, a2[j], b2[j]);
            if (act == 1) y = gn_silu2(y);
            const float2 f = gn_upk(y);
            o[2 * j] = __float_as_uint(f.x);
            o[2 * j + 1] = __float_as_uint(f.y);
        }
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// One 16-byte vector of the statistics pass: d = round_T(x + pre) - shift; s1 += d; s2 += d * d (per channel, packed pairs);
// nshift2 holds -shift.
template <typename T, bool HAS_PRE>
__device__ __forceinline__ void gn_stats_vec(const uint4& in, const uint64_t* nshift2, const uint64_t* pre2, uint64_t* s1, uint64_t* s2) {
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};
    if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint64_t x = GnPair<T>::unpack(w[j]);
            if (HAS_PRE) x = GnPair<T>::unpack(GnPair<T>::pack(gn_fadd2(x, pre2[j])));
            const uint64_t d = gn_fadd2(x, nshift2[j]);
            s1[j] = gn_fadd2(s1[j], d);
            s2[j] = gn_ffma2(d, d, s2[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint64_t x = gn_pk(__uint_as_float(w[2 * j]), __uint_as_float(w[2 * j + 1]));
            if (HAS_PRE) x = gn_fadd2(x, pre2[j]);
            const uint64_t d = gn_fadd2(x, nshift2[j]);
            s1[j] = gn_fadd2(s1[j], d);
            s2[j] = gn_ffma2(d, d, s2[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads, 2) gn_stats_kernel(const GnParams p) {
    constexpr int VN = GnVec<T>::N;
    extern __shared__ __align__(16) double sh_d[];                  // [2][Cs] channel sums (double), then float [2][R][Cs] + [Cs]
    double* ch_sum = sh_d;
    double* ch_sq = sh_d + p.Cs;
    float* st1 = reinterpret_cast<float*>(sh_d + 2 * p.Cs);   // [R][Cs] shifted sums per row-thread
    float* st2 = st1 + (size_t)p.R * p.Cs;                     // [R][Cs] shifted sums of squares
    float* stK = st2 + (size_t)p.R * p.Cs;                     // [Cs]    the shift of each channel
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int c_off = blockIdx.z * p.Cs;
    const int cv = threadIdx.x % p.CV, r = threadIdx.x / p.CV;
    const bool active = r < p.R;
    const T* xs = reinterpret_cast<const T*>(p.x) + (size_t)n * p.HW * p.C + c_off;
    const int row0 = chunk * p.rows_per_chunk;
    const int row1 = min(p.HW, row0 + p.rows_per_chunk);

    if (active) {
        constexpr int NP = VN / 2;
        float shift[VN], pre[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) pre[i] = 0.f;
        const bool has_pre = p.pre != nullptr;
        if (has_pre) gn_unpack<T>(*reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.pre) + (size_t)n * p.C + c_off + cv * VN), pre);
        gn_unpack<T>(*reinterpret_cast<const uint4*>(xs + (size_t)cv * VN), shift);   // pixel 0 of the sample
        uint64_t nshift2[NP], pre2[NP], s1[NP], s2[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            nshift2[j] = gn_pk(-Cvt<T>::round(shift[2 * j] + pre[2 * j]), -Cvt<T>::round(shift[2 * j + 1] + pre[2 * j + 1]));
            pre2[j] = gn_pk(pre[2 * j], pre[2 * j + 1]);
            s1[j] = 0ull;
            s2[j] = 0ull;
        }
        int row = row0 + r;
        // four independent 16-byte loads in flight per thread
        for (; row + 3 * p.R < row1; row += 4 * p.R) {
            uint4 in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) in[u] = ldg_stream16(xs + (size_t)(row + u * p.R) * p.C + cv * VN);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (has_pre) gn_stats_vec<T, true>(in[u], nshift2, pre2, s1, s2);
                else         gn_stats_vec<T, false>(in[u], nshift2, pre2, s1, s2);
            }
        }
        for (; row < row1; row += p.R) {
            const uint4 in = ldg_stream16(xs + (size_t)row * p.C + cv * VN);
            if (has_pre) gn_stats_vec<T, true>(in, nshift2, pre2, s1, s2);
            else         gn_stats_vec<T, false>(in, nshift2, pre2, s1, s2);
        }
        // park the per-thread shifted sums; the row-threads of a channel are combined below in the order r = 0..R-1
        float* d1 = st1 + (size_t)r * p.Cs + cv * VN;
        float* d2 = st2 + (size_t)r * p.Cs + cv * VN;
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
            const float2 a0 = gn_upk(s1[j]), a1 = gn_upk(s1[j + 1]), b0 = gn_upk(s2[j]), b1 = gn_upk(s2[j + 1]);
            *reinterpret_cast<float4*>(d1 + 2 * j) = make_float4(a0.x, a0.y, a1.x, a1.y);
            *reinterpret_cast<float4*>(d2 + 2 * j) = make_float4(b0.x, b0.y, b1.x, b1.y);
        }
        if (r == 0) {
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float2 k = gn_upk(nshift2[j]);
                stK[cv * VN + 2 * j] = -k.x;
                stK[cv * VN + 2 * j + 1] = -k.y;
            }
        }
    }
    __syncthreads();
    // un-shift in double and combine the R row-threads of each channel (fixed order => bit-reproducible)
    for (int c = threadIdx.x; c < p.Cs; c += blockDim.x) {
        const double K = (double)stK[c];
        double a = 0.0, b = 0.0;
        for (int rr = 0; rr < p.R; ++rr) {
            const int first = row0 + rr;
            const double nn = first < row1 ? (double)((row1 - first + p.R - 1) / p.R) : 0.0;
            const double d1 = (double)st1[(size_t)rr * p.Cs + c], d2 = (double)st2[(size_t)rr * p.Cs + c];
            a += d1 + nn * K;
            b += d2 + 2.0 * K * d1 + nn * K * K;
        }
        ch_sum[c] = a;
        ch_sq[c] = b;
    }
    __syncthreads();
    // channels -> groups (fixed order)
    const int Gs = p.Cs / p.cpg, g0 = c_off / p.cpg;
    for (int g = threadIdx.x; g < Gs; g += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int c = g * p.cpg; c < (g + 1) * p.cpg; ++c) { a += ch_sum[c]; b += ch_sq[c]; }
        double* out = p.partial + (((size_t)n * p.chunks + chunk) * p.G + g0 + g) * 2;
        out[0] = a;
        out[1] = b;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads, 2) gn_apply_kernel(const GnParams p) {
    constexpr int VN = GnVec<T>::N;
    extern __shared__ __align__(16) float sh_f[];                   // [2][Cs] : a_c, b_c
    __shared__ float g_mean[64], g_rstd[64];
    float* sa = sh_f;
    float* sb = sh_f + p.Cs;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int c_off = blockIdx.z * p.Cs;
    const int Gs = p.Cs / p.cpg, g0 = c_off / p.cpg;
    for (int g = threadIdx.x; g < Gs; g += blockDim.x) {
        double a = 0.0, b = 0.0;
        const double* part = p.partial + ((size_t)n * p.chunks * p.G + g0 + g) * 2;
        for (int c = 0; c < p.chunks; ++c) { a += part[(size_t)c * p.G * 2]; b += part[(size_t)c * p.G * 2 + 1]; }
        const double cnt = (double)p.HW * (double)p.cpg;
        const double mean = a / cnt;
        double var = b / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        g_mean[g] = (float)mean;
        g_rstd[g] = rsqrtf((float)var + p.eps);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.Cs; c += blockDim.x) {
        const int g = c / p.cpg;
        const float gam = p.gamma ? Cvt<T>::to_f(reinterpret_cast<const T*>(p.gamma)[c_off + c]) : 1.f;
        const float bet = p.beta ? Cvt<T>::to_f(reinterpret_cast<const T*>(p.beta)[c_off + c]) : 0.f;
        const float a = g_rstd[g] * gam;
        sa[c] = a;
        sb[c] = fmaf(-a, g_mean[g], bet);
    }
    __syncthreads();

    const int cv = threadIdx.x % p.CV, r = threadIdx.x / p.CV;
    if (r >= p.R) return;
    constexpr int NP = VN / 2;
    float pre[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) pre[i] = 0.f;
    const bool has_pre = p.pre != nullptr;
    if (has_pre) gn_unpack<T>(*reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.pre) + (size_t)n * p.C + c_off + cv * VN), pre);
    uint64_t a2[NP], b2[NP], pre2[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        a2[j] = gn_pk(sa[cv * VN + 2 * j], sa[cv * VN + 2 * j + 1]);
        b2[j] = gn_pk(sb[cv * VN + 2 * j], sb[cv * VN + 2 * j + 1]);
        pre2[j] = gn_pk(pre[2 * j], pre[2 * j + 1]);
    }
    const T* xs = reinterpret_cast<const T*>(p.x) + (size_t)n * p.HW * p.C + c_off;
    T* ys = reinterpret_cast<T*>(p.y) + (size_t)n * p.HW * p.C + c_off;
    const int row0 = chunk * p.rows_per_chunk;
    const int row1 = min(p.HW, row0 + p.rows_per_chunk);
    const int act = p.act;
    auto emit = [&](const uint4& in, size_t off) {
        *reinterpret_cast<uint4*>(ys + off) = has_pre ? gn_apply_vec<T, true>(in, a2, b2, pre2, act)
                                                      : gn_apply_vec<T, false>(in, a2, b2, pre2, act);
    };
    int row = row0 + r;
    // four independent 16-byte loads in flight per thread: with one, the 64 resident warps of an SM keep only 32 KB
    // in flight, ~3 TB/s across the chip at HBM latency
    for (; row + 3 * p.R < row1; row += 4 * p.R) {
        size_t off[4];
        uint4 in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            off[u] = (size_t)(row + u * p.R) * p.C + cv * VN;
            in[u] = ldg_stream16(xs + off[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) emit(in[u], off[u]);
    }
    for (; row < row1; row += p.R) {
        const size_t off = (size_t)row * p.C + cv * VN;
        emit(ldg_stream16(xs + off), off);
    }
}

// ---------------------------------------------------------------------------------------------
// y = x + bias[c] + vec[n][c] + residual  (one vectorised pass; replaces the separate bias / time-embedding /
// skip-connection element-wise kernels around the convolutions: resnet.py:341-376)
struct AddParams {
    const void* x; const void* bias; const void* vec; const void* res; void* y;
    int N, HW, C, CV, R, chunks, rows_per_chunk;
};

template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads) bias_residual_kernel(const AddParams p) {
    constexpr int VN = GnVec<T>::N;
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int CVT = blockDim.x / p.R;                 // column vectors handled per pass
    const int cv0 = threadIdx.x % CVT, r = threadIdx.x / CVT;
    const int row0 = chunk * p.rows_per_chunk;
    const int row1 = min(p.HW, row0 + p.rows_per_chunk);
    for (int cv = cv0; cv < p.CV; cv += CVT) {
        float add[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) add[i] = 0.f;
        if (p.bias) {
            float t[VN];
            gn_unpack<T>(*reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.bias) + cv * VN), t);
#pragma unroll
            for (int i = 0; i < VN; ++i) add[i] += t[i];
        }
        if (p.vec) {
            float t[VN];
            gn_unpack<T>(*reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.vec) + (size_t)n * p.C + cv * VN), t);
#pragma unroll
            for (int i = 0; i < VN; ++i) add[i] += t[i];
        }
        const size_t base = (size_t)n * p.HW * p.C + cv * VN;
        uint64_t add2[VN / 2];
#pragma unroll
        for (int j = 0; j < VN / 2; ++j) add2[j] = gn_pk(add[2 * j], add[2 * j + 1]);
        const bool has_res = p.res != nullptr;
        auto emit = [&](const uint4& xin, const uint4& rin, size_t off) {
            // (x + residual) + (bias + vec) in fp32, one rounding; packed pairs: same adds in half the issue slots
            const uint32_t xw[4] = {xin.x, xin.y, xin.z, xin.w}, rw[4] = {rin.x, rin.y, rin.z, rin.w};
            uint32_t o[4];
            if constexpr (sizeof(T) == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint64_t f = GnPair<T>::unpack(xw[j]);
                    if (has_res) f = gn_fadd2(f, GnPair<T>::unpack(rw[j]));
                    o[j] = GnPair<T>::pack(gn_fadd2(f, add2[j]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint64_t f = gn_pk(__uint_as_float(xw[2 * j]), __uint_as_float(xw[2 * j + 1]));
                    if (has_res) f = gn_fadd2(f, gn_pk(__uint_as_float(rw[2 * j]), __uint_as_float(rw[2 * j + 1])));
                    const float2 g = gn_upk(gn_fadd2(f, add2[j]));
                    o[2 * j] = __float_as_uint(g.x);
                    o[2 * j + 1] = __float_as_uint(g.y);
                }
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.y) + off) = make_uint4(o[0], o[1], o[2], o[3]);
        };
        int row = row0 + r;
        // four rows per trip: up to eight independent 16-byte loads in flight per thread before the first store
        for (; row + 3 * p.R < row1; row += 4 * p.R) {
            size_t off[4];
            uint4 xin[4], rin[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                off[u] = base + (size_t)(row + u * p.R) * p.C;
                xin[u] = ldg_stream16(reinterpret_cast<const T*>(p.x) + off[u]);
            }
            if (p.res) {
#pragma unroll
                for (int u = 0; u < 4; ++u) rin[u] = ldg_stream16(reinterpret_cast<const T*>(p.res) + off[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) emit(xin[u], rin[u], off[u]);
        }
        for (; row < row1; row += p.R) {
            const size_t off = base + (size_t)row * p.C;
            const uint4 xin = ldg_stream16(reinterpret_cast<const T*>(p.x) + off);
            uint4 rin = xin;
            if (p.res) rin = ldg_stream16(reinterpret_cast<const T*>(p.res) + off);
            emit(xin, rin, off);
        }
    }
}

// nearest-neighbour resize of an NHWC activation (upsampling.py Upsample2D: F.interpolate(mode="nearest"))
struct UpParams { const void* x; void* y; int N, Hi, Wi, Ho, Wo, CV; float sh, sw; };

template <typename T>
__global__ void __launch_bounds__(256) upsample_nearest_nhwc_kernel(const UpParams p) {
    const long long total = (long long)p.N * p.Ho * p.Wo * p.CV;
    const uint4* xin = reinterpret_cast<const uint4*>(p.x);
    uint4* yout = reinterpret_cast<uint4*>(p.y);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % p.CV);
        long long t = i / p.CV;
        const int wo = (int)(t % p.Wo); t /= p.Wo;
        const int ho = (int)(t % p.Ho);
        const int n = (int)(t / p.Ho);
        const int hi = min((int)floorf(ho * p.sh), p.Hi - 1);      // torch "nearest": floor(dst * in/out)
        const int wi = min((int)floorf(wo * p.sw), p.Wi - 1);
        yout[i] = __ldg(xin + (((long long)n * p.Hi + hi) * p.Wi + wi) * p.CV + cv);
    }
}

// Sub-pixel form of "nearest 2x upsampling followed by a 3x3 convolution" (upsampling.py Upsample2D.forward: F.interpolate
// (scale_factor=2, mode="nearest") then self.conv): output pixel (2i + a, 2j + b) only ever sees a 2x2 neighbourhood of the
// low-resolution input, so the convolution runs on the low-resolution tensor as ONE 2x2 convolution (padding 1) with
// 4 C output channels (channel block 2a + b = phase (a, b), tap weights = sums of the 3x3 taps that fall on the same source
// pixel; 16 instead of 36 multiply-adds per source pixel and output channel) and this kernel interleaves the four phases
// and adds the bias:   y[n, 2i + a, 2j + b, c] = t[n, i + a, j + b, (2a + b) C + c] + bias[c],   t = [N, H + 1, W + 1, 4 C].
struct PhaseParams { const void* t; const void* bias; void* y; int N, H, W, CV; };

template <typename T>
__global__ void __launch_bounds__(256) phase_shuffle_bias_kernel(const PhaseParams p) {
    constexpr int VN = GnVec<T>::N;
    constexpr int U = 4;                                   // independent 16-byte loads in flight per thread
    const int Ho = 2 * p.H, Wo = 2 * p.W, Hp = p.H + 1, Wp = p.W + 1;
    const long long total = (long long)p.N * Ho * Wo * p.CV;
    const T* tin = reinterpret_cast<const T*>(p.t);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    T* yout = reinterpret_cast<T*>(p.y);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
        uint4 in[U];
        int cvs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < total) {
                const int cv = (int)(i % p.CV);
                long long r = i / p.CV;
                const int wo = (int)(r % Wo); r /= Wo;
                const int ho = (int)(r % Ho);
                const int n = (int)(r / Ho);
                const int a = ho & 1, b = wo & 1;
                const long long src = ((((long long)n * Hp + (ho >> 1) + a) * Wp + (wo >> 1) + b) * 4 + (2 * a + b)) * p.CV + cv;
                in[u] = ldg_stream16(tin + src * VN);
                cvs[u] = cv;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < total) {
                if (bias) {
                    float f[VN], bv[VN];
                    gn_unpack<T>(in[u], f);
                    gn_unpack<T>(__ldg(reinterpret_cast<const uint4*>(bias) + cvs[u]), bv);
                    alignas(16) T o[VN];
#pragma unroll
                    for (int k = 0; k < VN; ++k) o[k] = Cvt<T>::from_f(f[k] + bv[k]);
                    *reinterpret_cast<uint4*>(yout + i * VN) = *reinterpret_cast<const uint4*>(o);
                } else {
                    *reinterpret_cast<uint4*>(yout + i * VN) = in[u];
                }
            }
        }
    }
}

// GEGLU gate of the transformer feed-forward (diffusers/src/diffusers/models/activations.py GEGLU.forward):
// proj [M, 2 D] -> out [M, D] = proj[:, :D] * gelu(proj[:, D:]); one pass instead of torch's gelu + mul over two
// strided halves.  Rounding as torch: gelu in fp32 -> T, product in fp32 -> T.
template <typename T>
__global__ void __launch_bounds__(256) geglu_kernel(const T* __restrict__ x, T* __restrict__ y, long long M, int DV) {
    constexpr int VN = GnVec<T>::N;
    const long long total = M * DV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / DV;
        const int dv = (int)(i - m * DV);
        const T* row = x + (size_t)m * (2 * DV * VN);
        float h[VN], g[VN];
        gn_unpack<T>(ldg_stream16(row + (size_t)dv * VN), h);
        gn_unpack<T>(ldg_stream16(row + (size_t)(DV + dv) * VN), g);
        alignas(16) T o[VN];
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            const float gel = Cvt<T>::round(g[k] * 0.5f * (1.0f + erff(g[k] * 0.70710678118654752440f)));
            o[k] = Cvt<T>::from_f(h[k] * gel);
        }
        *reinterpret_cast<uint4*>(y + (size_t)i * VN) = *reinterpret_cast<const uint4*>(o);
    }
}

// Row softmax with scale and cast for the VAE mid-block attention (one head of dim 512: its logits come out of a cuBLAS
// GEMM in fp32): out[r, :] = T(softmax(scale * logits[r, :])).  Replaces three element-wise passes over the [B, S, S]
// fp32 logits (scale in place, softmax, cast: 21 bytes of traffic per element) by one (4 read + 2 written).
// One 128-thread CTA per row at a time (grid-stride over rows), the row lives in registers: cols <= 128 * 4 * kSmMaxVec.
constexpr int kSmThreads = 128, kSmMaxVec = 20;
template <typename T>
__global__ void __launch_bounds__(kSmThreads) softmax_rows_kernel(const float* __restrict__ x, T* __restrict__ y, long long rows,
                                                                  int cols, float scale_log2) {
    __shared__ float sh[2][kSmThreads / 32];
    const int n4 = cols >> 2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const float4* xr = reinterpret_cast<const float4*>(x + r * cols);
        float4 v[kSmMaxVec];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kSmMaxVec; ++k) {
            const int i = tid + k * kSmThreads;
            if (i < n4) {
                const uint4 u = ldg_stream16(xr + i);
                v[k] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
                mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
            }
        }
        mx = warp_max(mx);
        if (lane == 0) sh[0][warp] = mx;
        __syncthreads();
        mx = fmaxf(fmaxf(sh[0][0], sh[0][1]), fmaxf(sh[0][2], sh[0][3]));
        // softmax(s x) = 2^((x - max) s log2 e) / sum for s > 0
        const float nb = -mx * scale_log2;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < kSmMaxVec; ++k) {
            const int i = tid + k * kSmThreads;
            if (i < n4) {
                v[k].x = exp2f(fmaf(v[k].x, scale_log2, nb)); v[k].y = exp2f(fmaf(v[k].y, scale_log2, nb));
                v[k].z = exp2f(fmaf(v[k].z, scale_log2, nb)); v[k].w = exp2f(fmaf(v[k].w, scale_log2, nb));
                sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
            }
        }
        sum = warp_sum(sum);
        if (lane == 0) sh[1][warp] = sum;
        __syncthreads();
        const float inv = __fdiv_rn(1.f, (sh[1][0] + sh[1][1]) + (sh[1][2] + sh[1][3]));
        T* yr = y + r * cols;
#pragma unroll
        for (int k = 0; k < kSmMaxVec; ++k) {
            const int i = tid + k * kSmThreads;
            if (i < n4) {
                alignas(8) T o[4] = {Cvt<T>::from_f(v[k].x * inv), Cvt<T>::from_f(v[k].y * inv), Cvt<T>::from_f(v[k].z * inv),
                                     Cvt<T>::from_f(v[k].w * inv)};
                if constexpr (sizeof(T) == 2) *reinterpret_cast<uint2*>(yr + 4 * i) = *reinterpret_cast<const uint2*>(o);
                else *reinterpret_cast<uint4*>(yr + 4 * i) = *reinterpret_cast<const uint4*>(o);
            }
        }
        __syncthreads();        // sh[] is reused by the next row
    }
}

// LayerNorm over the last axis of a contiguous [M, C] activation (the three norms of every BasicTransformerBlock,
// diffusers/src/diffusers/models/attention.py:447-448, 497-499, 520-521): one warp per row, the row lives in registers
// (two-pass mean / variance in fp32 like torch), y = T(((x - mean) * rstd) * gamma + beta).  One read + one write.
constexpr int kLnMaxVec = 8;          // 16-byte vectors per lane: C <= 32 * 8 * (16 / sizeof(T))
template <typename T>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ gamma,
                                                             const T* __restrict__ beta, long long M, int C, float eps) {
    constexpr int VN = GnVec<T>::N;
    const int nv = C / VN, lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const float inv_c = 1.f / (float)C;
    for (long long r = warp0; r < M; r += nwarps) {
        const T* xr = x + r * C;
        uint4 raw[kLnMaxVec];
#pragma unroll
        for (int k = 0; k < kLnMaxVec; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) raw[k] = ldg_stream16(xr + (size_t)i * VN);
        }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < kLnMaxVec; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) {
                float f[VN];
                gn_unpack<T>(raw[k], f);
#pragma unroll
                for (int q = 0; q < VN; ++q) sum += f[q];
            }
        }
        const float mean = warp_sum(sum) * inv_c;
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < kLnMaxVec; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) {
                float f[VN];
                gn_unpack<T>(raw[k], f);
#pragma unroll
                for (int q = 0; q < VN; ++q) { const float d = f[q] - mean; sq = fmaf(d, d, sq); }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) * inv_c + eps);
        T* yr = y + r * C;
#pragma unroll
        for (int k = 0; k < kLnMaxVec; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) {
                float f[VN], g[VN], b[VN];
                gn_unpack<T>(raw[k], f);
                gn_unpack<T>(*reinterpret_cast<const uint4*>(gamma + (size_t)i * VN), g);
                gn_unpack<T>(*reinterpret_cast<const uint4*>(beta + (size_t)i * VN), b);
                alignas(16) T o[VN];
#pragma unroll
                for (int q = 0; q < VN; ++q) o[q] = Cvt<T>::from_f(fmaf((f[q] - mean) * rstd, g[q], b[q]));
                *reinterpret_cast<uint4*>(yr + (size_t)i * VN) = *reinterpret_cast<const uint4*>(o);
            }
        }
    }
}

// Split HW rows into chunks (one CTA each, `per_chunk_ctas` CTAs per chunk index over samples / slices) so that the launch is
// whole waves of `slots` resident CTAs: minimises  waves x (rows per chunk + overhead_rows), chunks of at least min_rows rows.
static void plan_wave_chunks(int HW, long long per_chunk_ctas, int min_rows, int overhead_rows, int slots, int max_chunks,
                             int& chunks, int& rows) {
    int cmax = HW / (min_rows > 0 ? min_rows : 1);
    if (cmax < 1) cmax = 1;
    if (cmax > max_chunks) cmax = max_chunks;
    chunks = 1;
    rows = HW;
    double best = 1e300;
    for (int c = 1; c <= cmax; ++c) {
        const int rc = (HW + c - 1) / c;
        const int cc = (HW + rc - 1) / rc;                       // chunks actually needed with rc rows each
        const long long ctas = (long long)cc * per_chunk_ctas;
        const double waves = (double)((ctas + slots - 1) / slots);
        const double cost = waves * (double)(rc + overhead_rows);
        if (cost < best * (1.0 - 1e-9)) { best = cost; chunks = cc; rows = rc; }
    }
}

// channel slicing and CTA shape (needs C, G, cpg): Cs, CV, R; returns the slice count K or -1
static int gn_plan_slices(GnParams& p, int es) {
    const int VN = 16 / es;
    // slice the channels over blockIdx.z until one pixel row of a slice fits one CTA (<= 512 vectors);
    // slices must hold whole groups and whole 16-byte vectors
    int K = 1;
    for (; K <= 64; ++K) {
        if (p.C % K) continue;
        const int cs = p.C / K;
        if (cs % p.cpg == 0 && cs % VN == 0 && cs / VN <= kGnMaxThreads) break;
    }
    if (K > 64) return -1;
    p.Cs = p.C / K;
    p.CV = p.Cs / VN;
    int R = kGnMaxThreads / p.CV;
    if (R < 1) return -1;
    if (R > 32) R = 32;
    p.R = R;
    return K;
}

// + chunks per sample (needs N, HW and a current device)
static int gn_plan(GnParams& p, int es) {
    const int K = gn_plan_slices(p, es);
    if (K < 0) return -1;
    const int R = p.R;
    // Chunks per sample: two CTAs per SM are resident, so a launch runs in waves of 2 x SMs CTAs.  Pick the chunk count that
    // minimises  waves x (rows per chunk + a fixed per-CTA cost worth ~eight rows per thread: prologue that reduces the chunk
    // partials, first-load latency, reduction epilogue)  -- e.g. 12 samples -> 24 chunks = 288 CTAs = one full wave instead of 128
    // chunks = 5.2 waves of short CTAs (0.49 -> 0.69 of HBM on 12 x 108 x 192 x 512); 48 samples of 336 rows -> 6 chunks = 288 CTAs
    // instead of 7 = 1.13 waves.  At least 4 vectors per thread per chunk; per-CTA cost worth 8 rows per thread (2, 8 and 24 measured alike).
    int chunks, rows;
    plan_wave_chunks(p.HW, (long long)p.N * K, 4 * R, 8 * R, 2 * num_sms(), kGnMaxChunks, chunks, rows);
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

int rd_groupnorm_nhwc(const void* x, void* y, const void* gamma, const void* beta, const void* pre_add, int N, int HW,
                      int C, int G, float eps, int act, rd_dtype dtype, void* workspace, int64_t workspace_bytes,
                      void* stream) {
    RD_CHECK_ARG(x && y && workspace, "NULL pointer");
    RD_CHECK_ARG(N >= 1 && HW >= 1 && C >= 1 && G >= 1 && G <= 64 && C % G == 0, "bad shape N=%d HW=%d C=%d G=%d", N, HW, C, G);
    RD_CHECK_ARG(N <= 65535, "N must be <= 65535");
    RD_CHECK_ARG(act == 0 || act == 1, "act must be 0 (none) or 1 (SiLU)");
    RD_CHECK_ARG(dtype == RD_F32 || dtype == RD_F16 || dtype == RD_BF16, "bad dtype");
    const int es = dtype == RD_F32 ? 4 : 2;
    if (C % (16 / es) != 0) { set_error("C=%d is not a multiple of the 16-byte vector width", C); return RD_ERR_UNSUPPORTED; }
    RD_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)pre_add & 15) == 0,
                 "x / y / pre_add must be 16-byte aligned");
    GnParams p{};
    p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.pre = pre_add;
    p.partial = reinterpret_cast<double*>(workspace);
    p.N = N; p.HW = HW; p.C = C; p.G = G; p.cpg = C / G; p.eps = eps; p.act = act;
    if (gn_plan(p, es) != 0) { set_error("no channel slicing found for C=%d G=%d", C, G); return RD_ERR_UNSUPPORTED; }
    if (workspace_bytes < rd_groupnorm_workspace_bytes(N, G)) {
        set_error("group-norm workspace too small");
        return RD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = p.CV * p.R;
    dim3 grid(p.chunks, N, C / p.Cs);
    // (measured and not kept -- tools/probes/gn_variants, profiles/r2_gn_variants.txt, profiles/r2_gn_staged.txt: a per-thread
    //  cp.async staging ring, 3-6 % slower; a producer-warp + bulk-copy ring pair with a finalise kernel, no better than these two
    //  kernels once the chunks are wave-sized; an apply pass of many short CTAs, -1...-10 %)
    const size_t sm_stats = 2 * (size_t)p.Cs * sizeof(double) + (2 * (size_t)p.R + 1) * p.Cs * sizeof(float);
    const size_t sm_apply = 2 * (size_t)p.Cs * sizeof(float);
    {   // up to 16 B * Cs + 8 B * R * Cs = 96 KB for the widest slice: opt in above the 48 KB default (per device)
        static PerDeviceOnce configured;
        if (configured.needed()) {
            RD_CUDA(cudaFuncSetAttribute(gn_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            RD_CUDA(cudaFuncSetAttribute(gn_stats_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            RD_CUDA(cudaFuncSetAttribute(gn_stats_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            configured.done();
        }
    }
    switch (dtype) {
        case RD_F32:
            gn_stats_kernel<float><<<grid, threads, sm_stats, st>>>(p);
            gn_apply_kernel<float><<<grid, threads, sm_apply, st>>>(p);
            break;
        case RD_F16:
            gn_stats_kernel<__half><<<grid, threads, sm_stats, st>>>(p);
            gn_apply_kernel<__half><<<grid, threads, sm_apply, st>>>(p);
            break;
        default:
            gn_stats_kernel<__nv_bfloat16><<<grid, threads, sm_stats, st>>>(p);
            gn_apply_kernel<__nv_bfloat16><<<grid, threads, sm_apply, st>>>(p);
            break;
    }
    RD_LAUNCH_CHECK("gn kernels");
    return RD_OK;
}

int rd_bias_residual_nhwc(const void* x, const void* bias, const void* vec, const void* residual, void* y, int N, int HW,
                          int C, rd_dtype dtype, void* stream) {
    RD_CHECK_ARG(x && y, "NULL pointer");
    RD_CHECK_ARG(N >= 1 && N <= 65535 && HW >= 1 && C >= 1, "bad shape");
    RD_CHECK_ARG(dtype == RD_F32 || dtype == RD_F16 || dtype == RD_BF16, "bad dtype");
    const int es = dtype == RD_F32 ? 4 : 2, VN = 16 / es;
    if (C % VN != 0) { set_error("C=%d is not a multiple of the 16-byte vector width", C); return RD_ERR_UNSUPPORTED; }
    RD_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)vec | (uintptr_t)residual) & 15) == 0,
                 "pointers must be 16-byte aligned");
    AddParams p{};
    p.x = x; p.bias = bias; p.vec = vec; p.res = residual; p.y = y; p.N = N; p.HW = HW; p.C = C; p.CV = C / VN;
    int cvt = p.CV < 256 ? p.CV : 256;          // column vectors per pass
    while (p.CV % cvt) --cvt;
    int R = 256 / cvt; if (R < 1) R = 1; if (R > 32) R = 32;
    p.R = R;
    // many short CTAs (16 rows per thread: four trips of four rows).  One wave of long CTAs, which is what GroupNorm wants
    // (plan_wave_chunks), was measured SLOWER here: 0.87 / 0.85 / 0.78 of HBM against 0.98 / 0.92 / 0.83 on the decoder levels --
    // this kernel has no per-CTA prologue or reduction to amortise, and short CTAs balance the DRAM channels dynamically.
    int rows = 16 * R;
    int chunks = (HW + rows - 1) / rows;
    if (chunks > 4096) { chunks = 4096; rows = (HW + chunks - 1) / chunks; chunks = (HW + rows - 1) / rows; }
    p.chunks = chunks; p.rows_per_chunk = rows;
    dim3 grid(chunks, N);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case RD_F32: bias_residual_kernel<float><<<grid, cvt * R, 0, st>>>(p); break;
        case RD_F16: bias_residual_kernel<__half><<<grid, cvt * R, 0, st>>>(p); break;
        default:     bias_residual_kernel<__nv_bfloat16><<<grid, cvt * R, 0, st>>>(p); break;
    }
    RD_LAUNCH_CHECK("bias_residual_kernel");
    return RD_OK;
}

int rd_upsample_nearest_nhwc(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, rd_dtype dtype,
                             void* stream) {
    RD_CHECK_ARG(x && y && N >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1 && C >= 1, "bad argument");
    RD_CHECK_ARG(dtype == RD_F32 || dtype == RD_F16 || dtype == RD_BF16, "bad dtype");
    const int es = dtype == RD_F32 ? 4 : 2, VN = 16 / es;
    if (C % VN != 0) { set_error("C=%d is not a multiple of the 16-byte vector width", C); return RD_ERR_UNSUPPORTED; }
    RD_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "pointers must be 16-byte aligned");
    UpParams p{x, y, N, Hi, Wi, Ho, Wo, C / VN, (float)Hi / (float)Ho, (float)Wi / (float)Wo};
    const long long total = (long long)N * Ho * Wo * p.CV;
    long long want = (total + 255) / 256;
    const int blocks = (int)(want > (long long)num_sms() * 32 ? (long long)num_sms() * 32 : want);
    upsample_nearest_nhwc_kernel<uint4><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    RD_LAUNCH_CHECK("upsample_nearest_nhwc_kernel");
    return RD_OK;
}

int rd_phase_shuffle_bias_nhwc(const void* t, const void* bias, void* y, int N, int H, int W, int C, rd_dtype dtype,
                               void* stream) {
    RD_CHECK_ARG(t && y && N >= 1 && H >= 1 && W >= 1 && C >= 1, "bad argument");
    RD_CHECK_ARG(dtype == RD_F32 || dtype == RD_F16 || dtype == RD_BF16, "bad dtype");
    const int es = dtype == RD_F32 ? 4 : 2, VN = 16 / es;
    if (C % VN != 0) { set_error("C=%d is not a multiple of the 16-byte vector width", C); return RD_ERR_UNSUPPORTED; }
    RD_CHECK_ARG((((uintptr_t)t | (uintptr_t)y | (uintptr_t)bias) & 15) == 0, "pointers must be 16-byte aligned");
    PhaseParams p{t, bias, y, N, H, W, C / VN};
    const long long total = (long long)N * 4 * H * W * p.CV;
    const long long want = (total + 4 * 256 - 1) / (4 * 256);
    const int blocks = (int)(want > (long long)num_sms() * 16 ? (long long)num_sms() * 16 : want);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case RD_F32: phase_shuffle_bias_kernel<float><<<blocks, 256, 0, st>>>(p); break;
        case RD_F16: phase_shuffle_bias_kernel<__half><<<blocks, 256, 0, st>>>(p); break;
        default:     phase_shuffle_bias_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p); break;
    }
    RD_LAUNCH_CHECK("phase_shuffle_bias_kernel");
    return RD_OK;
}

int rd_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t M, int C, float eps, rd_dtype dtype,
                 void* stream) {
    RD_CHECK_ARG(x && y && gamma && beta && M >= 1 && C >= 1, "bad argument");
    RD_CHECK_ARG(dtype == RD_F32 || dtype == RD_F16 || dtype == RD_BF16, "bad dtype");
    const int es = dtype == RD_F32 ? 4 : 2, VN = 16 / es;
    if (C % VN != 0 || C > 32 * kLnMaxVec * VN) {
        set_error("C=%d: rd_layernorm needs a multiple of %d and at most %d channels", C, VN, 32 * kLnMaxVec * VN);
        return RD_ERR_UNSUPPORTED;
    }
    RD_CHECK_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0, "pointers must be 16-byte aligned");
    const long long want = (M + 7) / 8;
    const int blocks = (int)(want > (long long)num_sms() * 8 ? (long long)num_sms() * 8 : want);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case RD_F32: layernorm_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, (const float*)gamma, (const float*)beta, M, C, eps); break;
        case RD_F16: layernorm_rows_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)x, (__half*)y, (const __half*)gamma, (const __half*)beta, M, C, eps); break;
        default:     layernorm_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, M, C, eps); break;
    }
    RD_LAUNCH_CHECK("layernorm_rows_kernel");
    return RD_OK;
}

int rd_softmax_rows(const float* logits, void* out, int64_t rows, int cols, float scale, rd_dtype out_dtype, void* stream) {
    RD_CHECK_ARG(logits && out && rows >= 1 && cols >= 1, "bad argument");
    RD_CHECK_ARG(scale > 0.f && isfinite(scale), "scale must be positive and finite");
    RD_CHECK_ARG(out_dtype == RD_F32 || out_dtype == RD_F16 || out_dtype == RD_BF16, "bad dtype");
    if (cols % 4 != 0 || cols > 4 * kSmThreads * kSmMaxVec) {
        set_error("cols=%d: rd_softmax_rows needs a multiple of 4 and at most %d columns", cols, 4 * kSmThreads * kSmMaxVec);
        return RD_ERR_UNSUPPORTED;
    }
    RD_CHECK_ARG((((uintptr_t)logits | (uintptr_t)out) & 15) == 0, "pointers must be 16-byte aligned");
    const long long want = rows;
    const int blocks = (int)(want > (long long)num_sms() * 16 ? (long long)num_sms() * 16 : want);
    const float sl2 = scale * 1.4426950408889634f;
    cudaStream_t st = (cudaStream_t)stream;
    switch (out_dtype) {
        case RD_F32: softmax_rows_kernel<float><<<blocks, kSmThreads, 0, st>>>(logits, (float*)out, rows, cols, sl2); break;
        case RD_F16: softmax_rows_kernel<__half><<<blocks, kSmThreads, 0, st>>>(logits, (__half*)out, rows, cols, sl2); break;
        default:     softmax_rows_kernel<__nv_bfloat16><<<blocks, kSmThreads, 0, st>>>(logits, (__nv_bfloat16*)out, rows, cols, sl2); break;
    }
    RD_LAUNCH_CHECK("softmax_rows_kernel");
    return RD_OK;
}

int rd_geglu(const void* proj, void* y, int64_t M, int D, rd_dtype dtype, void* stream) {
    RD_CHECK_ARG(proj && y && M >= 1 && D >= 1, "bad argument");
    RD_CHECK_ARG(dtype == RD_F32 || dtype == RD_F16 || dtype == RD_BF16, "bad dtype");
    const int es = dtype == RD_F32 ? 4 : 2, VN = 16 / es;
    if (D % VN != 0) { set_error("D=%d is not a multiple of the 16-byte vector width", D); return RD_ERR_UNSUPPORTED; }
    RD_CHECK_ARG((((uintptr_t)proj | (uintptr_t)y) & 15) == 0, "pointers must be 16-byte aligned");
    const int DV = D / VN;
    const long long total = (long long)M * DV;
    long long want = (total + 255) / 256;
    const int blocks = (int)(want > (long long)num_sms() * 16 ? (long long)num_sms() * 16 : want);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case RD_F32: geglu_kernel<float><<<blocks, 256, 0, st>>>((const float*)proj, (float*)y, M, DV); break;
        case RD_F16: geglu_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)proj, (__half*)y, M, DV); break;
        default:     geglu_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)proj, (__nv_bfloat16*)y, M, DV); break;
    }
    RD_LAUNCH_CHECK("geglu_kernel");
    return RD_OK;
}

}  // extern "C"
