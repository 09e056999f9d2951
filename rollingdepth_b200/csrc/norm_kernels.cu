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
    int staged;            // 1: loads through the cp.async staging ring (gn_stream_rows), 0: four register loads per trip
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

// ---- per-thread staging ring filled by cp.async (LDGSTS): loads in flight cost shared memory, not registers.
// Slot s of thread t is ring[s * blockDim.x + t] (16 bytes, consecutive threads consecutive: conflict-free).  Two groups
// of kGnGroup rows: while one group is being consumed the other one is in flight, so a thread keeps 4-8 16-byte loads
// outstanding and an SM with 1024 resident threads 64-128 KB -- the streaming kernels of this file went from 0.72 to
// 0.98 of the HBM copy bandwidth when their loads in flight per SM doubled from 64 KB to 128 KB (bias_residual).
constexpr int kGnGroup = 4, kGnSlots = 2 * kGnGroup;
__device__ __forceinline__ void gn_cp_async16(uint4* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void gn_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gn_cp_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// rows row_first, row_first + R, ... < row1 of one 16-byte column: `consume(vector, row)` for each, in row order
template <typename T, typename F>
__device__ __forceinline__ void gn_stream_rows(const T* col, size_t row_stride, int row_first, int row1, int R, uint4* ring,
                                               F&& consume) {
    const int T_ = blockDim.x, tid = threadIdx.x;
    const int nrows = row_first < row1 ? (row1 - row_first + R - 1) / R : 0;
    const int ntrips = (nrows + kGnGroup - 1) / kGnGroup;
    auto issue = [&](int trip) {
        const int grp = trip & 1;
#pragma unroll
        for (int u = 0; u < kGnGroup; ++u) {
            const int k = trip * kGnGroup + u;
            if (k < nrows) gn_cp_async16(ring + (size_t)(grp * kGnGroup + u) * T_ + tid, col + (size_t)(row_first + k * R) * row_stride);
        }
        gn_cp_commit();                     // also when empty: one group per call keeps the wait accounting uniform
    };
    issue(0);
    issue(1);
    for (int trip = 0; trip < ntrips; ++trip) {
        gn_cp_wait_but_one();               // groups complete in order: the group of `trip` has landed
        const int grp = trip & 1;
        uint4 v[kGnGroup];
#pragma unroll
        for (int u = 0; u < kGnGroup; ++u) v[u] = ring[(size_t)(grp * kGnGroup + u) * T_ + tid];
        issue(trip + 2);                    // refill the slots just read
#pragma unroll
        for (int u = 0; u < kGnGroup; ++u) {
            const int k = trip * kGnGroup + u;
            if (k < nrows) consume(v[u], row_first + k * R);
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

    float shift[VN], s1[VN], s2[VN], pre[VN];
    if (active) {
#pragma unroll
        for (int i = 0; i < VN; ++i) pre[i] = 0.f;
        if (p.pre) gn_unpack<T>(*reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.pre) + (size_t)n * p.C + c_off + cv * VN), pre);
        gn_unpack<T>(*reinterpret_cast<const uint4*>(xs + (size_t)cv * VN), shift);   // pixel 0 of the sample
#pragma unroll
        for (int i = 0; i < VN; ++i) { shift[i] = Cvt<T>::round(shift[i] + pre[i]); s1[i] = 0.f; s2[i] = 0.f; }
        int row = row0 + r;
        if (p.staged) {
            uint4* ring = reinterpret_cast<uint4*>(stK + p.Cs);
            gn_stream_rows<T>(xs + (size_t)cv * VN, (size_t)p.C, row, row1, p.R, ring, [&](const uint4& in, int) {
                float fa[VN];
                gn_unpack<T>(in, fa);
#pragma unroll
                for (int i = 0; i < VN; ++i) {
                    const float da = Cvt<T>::round(fa[i] + pre[i]) - shift[i];
                    s1[i] += da;
                    s2[i] = fmaf(da, da, s2[i]);
                }
            });
            row = row1;
        }
        // four independent 16-byte loads in flight per thread
        for (; row + 3 * p.R < row1; row += 4 * p.R) {
            uint4 in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) in[u] = ldg_stream16(xs + (size_t)(row + u * p.R) * p.C + cv * VN);
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
                float fa[VN], fb[VN];
                gn_unpack<T>(in[u], fa);
                gn_unpack<T>(in[u + 1], fb);
#pragma unroll
                for (int i = 0; i < VN; ++i) {
                    const float da = Cvt<T>::round(fa[i] + pre[i]) - shift[i], db = Cvt<T>::round(fb[i] + pre[i]) - shift[i];
                    s1[i] += da + db;
                    s2[i] = fmaf(da, da, fmaf(db, db, s2[i]));
                }
            }
        }
        for (; row < row1; row += p.R) {
            float fa[VN];
            gn_unpack<T>(ldg_stream16(xs + (size_t)row * p.C + cv * VN), fa);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
                const float da = Cvt<T>::round(fa[i] + pre[i]) - shift[i];
                s1[i] += da;
                s2[i] = fmaf(da, da, s2[i]);
            }
        }
        // park the per-thread shifted sums; the row-threads of a channel are combined below in the order r = 0..R-1
        float* d1 = st1 + (size_t)r * p.Cs + cv * VN;
        float* d2 = st2 + (size_t)r * p.Cs + cv * VN;
#pragma unroll
        for (int i = 0; i < VN; i += 4) {
            *reinterpret_cast<float4*>(d1 + i) = make_float4(s1[i], s1[i + 1], s1[i + 2], s1[i + 3]);
            *reinterpret_cast<float4*>(d2 + i) = make_float4(s2[i], s2[i + 1], s2[i + 2], s2[i + 3]);
        }
        if (r == 0) {
#pragma unroll
            for (int i = 0; i < VN; ++i) stK[cv * VN + i] = shift[i];
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
    float a[VN], b[VN], pre[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) { a[i] = sa[cv * VN + i]; b[i] = sb[cv * VN + i]; pre[i] = 0.f; }
    if (p.pre) gn_unpack<T>(*reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.pre) + (size_t)n * p.C + c_off + cv * VN), pre);
    const T* xs = reinterpret_cast<const T*>(p.x) + (size_t)n * p.HW * p.C + c_off;
    T* ys = reinterpret_cast<T*>(p.y) + (size_t)n * p.HW * p.C + c_off;
    const int row0 = chunk * p.rows_per_chunk;
    const int row1 = min(p.HW, row0 + p.rows_per_chunk);
    auto emit = [&](const uint4& in, size_t off) {
        float f[VN];
        gn_unpack<T>(in, f);
        alignas(16) T o[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) {
            const float xin = p.pre ? Cvt<T>::round(f[i] + pre[i]) : f[i];
            float v = Cvt<T>::round(fmaf(xin, a[i], b[i]));        // group_norm output, rounded to T like torch
            if (p.act == 1) v = __fdividef(v, 1.f + __expf(-v));    // SiLU on the rounded value
            o[i] = Cvt<T>::from_f(v);
        }
        *reinterpret_cast<uint4*>(ys + off) = *reinterpret_cast<const uint4*>(o);
    };
    int row = row0 + r;
    if (p.staged) {
        uint4* ring = reinterpret_cast<uint4*>(sh_f + 2 * p.Cs);
        gn_stream_rows<T>(xs + (size_t)cv * VN, (size_t)p.C, row, row1, p.R, ring,
                          [&](const uint4& in, int rw) { emit(in, (size_t)rw * p.C + cv * VN); });
        return;
    }
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
        auto emit = [&](const uint4& xin, const uint4& rin, size_t off) {
            float f[VN];
            gn_unpack<T>(xin, f);
            if (p.res) {
                float g[VN];
                gn_unpack<T>(rin, g);
#pragma unroll
                for (int i = 0; i < VN; ++i) f[i] += g[i];
            }
            alignas(16) T o[VN];
#pragma unroll
            for (int i = 0; i < VN; ++i) o[i] = Cvt<T>::from_f(f[i] + add[i]);
            *reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.y) + off) = *reinterpret_cast<const uint4*>(o);
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

static int gn_plan(GnParams& p, int es) {
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
    if (workspace_bytes < (int64_t)N * p.chunks * G * 2 * (int64_t)sizeof(double)) {
        set_error("group-norm workspace too small");
        return RD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = p.CV * p.R;
    dim3 grid(p.chunks, N, C / p.Cs);
    static int staged_env = -1;
    if (staged_env < 0) { const char* e = getenv("RD_GN_STAGED"); staged_env = e ? (atoi(e) != 0) : 1; }
    p.staged = staged_env;
    const size_t ring_bytes = p.staged ? (size_t)kGnSlots * 16 * threads : 0;
    // the ring starts 16-byte aligned: (2 Cs doubles + (2 R + 1) Cs floats) and (2 Cs floats) with Cs % 4 == 0
    const size_t sm_stats = 2 * (size_t)p.Cs * sizeof(double) + (2 * (size_t)p.R + 1) * p.Cs * sizeof(float) + ring_bytes;
    const size_t sm_apply = 2 * (size_t)p.Cs * sizeof(float) + ring_bytes;
    {   // up to 16 B * Cs + 8 B * R * Cs = 96 KB for the widest slice + 64 KB staging ring: opt in above the 48 KB default
        static PerDeviceOnce configured;
        if (configured.needed()) {
            RD_CUDA(cudaFuncSetAttribute(gn_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024));
            RD_CUDA(cudaFuncSetAttribute(gn_stats_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024));
            RD_CUDA(cudaFuncSetAttribute(gn_stats_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024));
            RD_CUDA(cudaFuncSetAttribute(gn_apply_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            RD_CUDA(cudaFuncSetAttribute(gn_apply_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            RD_CUDA(cudaFuncSetAttribute(gn_apply_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
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
    int rows = 16 * R;                    // 16 rows per thread: four trips of four rows
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
