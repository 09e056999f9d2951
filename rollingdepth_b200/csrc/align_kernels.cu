// Depth co-alignment kernels (sm_100a): global min, crop/subsample, the fused per-iteration
// optimisation pass (closed-form gradient + Adam), overlap-average merge, renormalisation.
//
// Replaces rollingdepth/depth_aligner.py:68-262 and rollingdepth_pipeline.py:316-318 of the
// reference.  All of these are HBM-bound streaming kernels: coalesced vector loads, warp-shuffle
// reductions, no atomics on floats (results are deterministic run to run).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace rd {

// =============================================================================================
// Generic "block partials -> last block reduces" helper used by min / minmax reductions
// =============================================================================================
constexpr int kRedMaxBlocks = 16384;
struct RedScratch {
    float part_min[kRedMaxBlocks];
    float part_max[kRedMaxBlocks];
    unsigned int counter;
};

struct SegList {
    const void* ptr[RD_MAX_GROUPS];
    long long count[RD_MAX_GROUPS];
    int nseg;
};

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void unpack16(const uint4& u, float* out) {
    if constexpr (sizeof(T) == 4) {
        out[0] = __uint_as_float(u.x); out[1] = __uint_as_float(u.y);
        out[2] = __uint_as_float(u.z); out[3] = __uint_as_float(u.w);
    } else {
        const T* h = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = Cvt<T>::to_f(h[i]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) global_min_kernel(SegList segs, RedScratch* sc, float* out_min) {
    float mn = INFINITY;
    constexpr int VN = Vec16<T>::N;
    for (int sgi = 0; sgi < segs.nseg; ++sgi) {
        const T* base = reinterpret_cast<const T*>(segs.ptr[sgi]);
        const long long n = segs.count[sgi];
        // head elements until 16-byte alignment
        long long head = ((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15) / sizeof(T);
        if (head > n) head = n;
        const long long nvec = (n - head) / VN;
        const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long gstride = (long long)gridDim.x * blockDim.x;
        if (gtid < head) mn = fminf(mn, Cvt<T>::to_f(base[gtid]));
        const uint4* vp = reinterpret_cast<const uint4*>(base + head);
        long long i = gtid;
        // 4 independent 16-byte loads in flight per thread
        for (; i + 3 * gstride < nvec; i += 4 * gstride) {
            uint4 a = ldg_stream16(vp + i), b = ldg_stream16(vp + i + gstride);
            uint4 c = ldg_stream16(vp + i + 2 * gstride), d = ldg_stream16(vp + i + 3 * gstride);
            float f[VN];
            unpack16<T>(a, f);
#pragma unroll
            for (int q = 0; q < VN; ++q) mn = fminf(mn, f[q]);
            unpack16<T>(b, f);
#pragma unroll
            for (int q = 0; q < VN; ++q) mn = fminf(mn, f[q]);
            unpack16<T>(c, f);
#pragma unroll
            for (int q = 0; q < VN; ++q) mn = fminf(mn, f[q]);
            unpack16<T>(d, f);
#pragma unroll
            for (int q = 0; q < VN; ++q) mn = fminf(mn, f[q]);
        }
        for (; i < nvec; i += gstride) {
            float f[VN];
            unpack16<T>(ldg_stream16(vp + i), f);
#pragma unroll
            for (int q = 0; q < VN; ++q) mn = fminf(mn, f[q]);
        }
        const long long tail0 = head + nvec * VN;
        if (tail0 + gtid < n) mn = fminf(mn, Cvt<T>::to_f(base[tail0 + gtid]));
    }
    __shared__ float sh[8];
    __shared__ bool last;
    mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = sh[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fminf(m, sh[w]);
        sc->part_min[blockIdx.x] = m;
        __threadfence();
        unsigned prev = atomicAdd(&sc->counter, 1u);
        last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        float m = INFINITY;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) m = fminf(m, __ldcg(&sc->part_min[i]));
        m = warp_min(m);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float r = sh[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fminf(r, sh[w]);
            out_min[0] = r;
            sc->counter = 0;
        }
    }
}

// =============================================================================================
// prep: (x - min) in the snippet dtype, border crop, stride-`factor` subsample
// =============================================================================================
struct PrepParams {
    const void* snip[RD_MAX_GROUPS];
    rd_coalign_problem pb;
};

template <typename T>
__global__ void __launch_bounds__(256) prep_kernel(PrepParams pp, const float* dev_min, T* out) {
    const rd_coalign_problem& pb = pp.pb;
    const long long total = (long long)pb.n_total * pb.w * pb.P_pad;
    const float mn = dev_min[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % pb.P_pad);
        const long long rj = i / pb.P_pad;  // row * w + j
        float v = 0.f;
        if (p < pb.P) {
            const int row = (int)(rj / pb.w);
            int g = 0;
#pragma unroll
            for (int q = 1; q < RD_MAX_GROUPS; ++q)
                if (q < pb.G && row >= pb.row0[q]) g = q;
            const long long local = rj - (long long)pb.row0[g] * pb.w;  // k * w + j inside group g
            const int hs = p / pb.Ws, ws = p % pb.Ws;
            const int h = pb.border + hs * pb.factor, w = pb.border + ws * pb.factor;
            const T* src = reinterpret_cast<const T*>(pp.snip[g]);
            const float x = Cvt<T>::to_f(src[(local * pb.H + h) * pb.W + w]);
            v = Cvt<T>::round(__fsub_rn(x, mn));  // torch: arithmetic in fp32, rounded once to the tensor dtype
        }
        out[i] = Cvt<T>::from_f(v);
    }
}

// =============================================================================================
// The fused optimisation pass
// =============================================================================================
// CTA shape of the pass: the packed accumulators need ~168 registers per thread, i.e. <= 384 threads per SM.
// shape 0: 192 threads x 2 CTAs/SM, 96 KB staged each (best while the subsampled snippets stay L2-resident);
// shape 1: 128 threads x 3 CTAs/SM, 64 KB staged each (three-way overlap of the load / compute / reduce phases of
//          consecutive CTAs: best when every iteration streams from HBM).  Chosen per problem by its working set;
//          RD_COALIGN_CFG=0/1 forces one of them (A/B measurements).
constexpr int kPassV = 2;                 // pixels per thread per step (one 4- or 8-byte shared-memory load)
constexpr int kPassSmemBudgetMax = 98304;
constexpr int kFrameStats = 8;            // sum|target|, sum|tdep|, L1, L2, tmin, tmax, pad, pad

struct PassParams {
    const void* x;
    int G, w, N, P, P_pad, R, C, n_total, px_per_chunk;
    int d[RD_MAX_GROUPS], n[RD_MAX_GROUPS], row0[RD_MAX_GROUPS];
    int f_begin;
    int pdl;                 // host-side only: launch with programmatic stream serialization (see launch_pass_kernel)
    int apply_update;
    float step_size, bc2_sqrt, beta1, beta2, eps;
    float lmda2, lmda3, dlw, ls_over_Z, inv_P;
    const float* theta_in;   // [6][n_total] : s, t, m_s, m_t, v_s, v_t
    float* theta_out;
    const float* gterm_in;   // [N][R][2] from the previous pass (after exchange)
    float* gterm_out;        // [N][R][2] written by this pass
    float* part;             // [N][C][4R + kFrameStats]
    unsigned int* counter;   // [N]
    float* hist_frame;       // [N][4] for this pass: loss term, soft term, tmin, tmax
};

__device__ __forceinline__ void adam_update(float& theta, float& m, float& v, float g, const PassParams& p) {
    // torch.optim.Adam single-tensor path: lerp for m, mul/addcmul for v, addcdiv for theta
    m = m + (g - m) * (1.f - p.beta1);
    v = v * p.beta2 + (1.f - p.beta2) * g * g;
    const float denom = __fdiv_rn(sqrtf(v), p.bc2_sqrt) + p.eps;
    theta = theta - __fdiv_rn(p.step_size * m, denom);
}

__device__ __forceinline__ uint32_t pass_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 1-D bulk asynchronous copy global -> shared (TMA engine, no registers, completion on an mbarrier)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pass_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(pass_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU.RCP, <= 1 ulp
    return y;
}
// x with the sign of r (r = +0 counts as positive; exact zeros only occur for single-contribution frames,
// which are handled separately)
__device__ __forceinline__ float with_sign_of(float x, float r) {
    return __int_as_float(__float_as_int(x) ^ (__float_as_int(r) & 0x80000000));
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 take one issue slot for two pixels; a {s, s} operand
// is encoded as a broadcast scalar register).  The pass is issue-bound, so the two adjacent pixels of a step travel as
// one 64-bit register pair through every element-wise operation and through the accumulators.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t bc2(float s) { return pk2(s, s); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// two adjacent pixels of one staged row as a packed fp32 pair
template <typename T>
__device__ __forceinline__ uint64_t lds_pair(const T* p) {
    if constexpr (sizeof(T) == 4) {
        return *reinterpret_cast<const uint64_t*>(p);
    } else {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
        const T* h = reinterpret_cast<const T*>(&u);
        return pk2(Cvt<T>::to_f(h[0]), Cvt<T>::to_f(h[1]));
    }
}
// (mag | sign bit of r) per lane: mag = 1.0f gives sign(r) with sign(+0) = +1; mag = 0.0f gives +-0
__device__ __forceinline__ uint64_t signed_mag2(float mag_lo, float mag_hi, uint64_t r) {
    float r0, r1;
    upk2(r, r0, r1);
    return pk2(__int_as_float((__float_as_int(r0) & 0x80000000) | __float_as_int(mag_lo)),
               __int_as_float((__float_as_int(r1) & 0x80000000) | __float_as_int(mag_hi)));
}

// Sum NVAL per-thread values over the 32 lanes of a warp with 31 shuffles per 32 values (butterfly transpose):
// afterwards lane L holds the warp total of value index bitrev-free mapping idx(L) = L for the first 32 slots.
template <int NSLOT>   // NSLOT = 32
__device__ __forceinline__ float warp_transpose_sum(float (&a)[NSLOT], int lane) {
#pragma unroll
    for (int k = NSLOT / 2; k >= 1; k >>= 1) {
        const bool hi = (lane & k) != 0;
#pragma unroll
        for (int i = 0; i < k; ++i) {
            const float send = hi ? a[i] : a[i + k];
            const float keep = hi ? a[i + k] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, k);
        }
    }
    return a[0];   // total of slot `lane` (slot index bits are consumed MSB-first, matching the lane bits)
}

template <int RM>
struct PassAcc {                 // every member holds the partial sums of the even and of the odd pixel
    uint64_t A0[RM], A1[RM], D0[RM], D1[RM];
    uint64_t fa, fb, l1;
    float l2, tmin, tmax;
};

// One step = two adjacent pixels (q0, q0 + 1) of every contribution.  ALLV: all RM contributions are present
// (no predicates); TAIL: only pixel q0 is real (the odd last pixel of a frame; its neighbour is row padding).
// ~13 issue slots per (contribution, pixel): LDS.64, FFMA2, 2 FMNMX, 2 MUFU.RCP, 2 FADD2 | FFMA2, FSUB2, 2 LOP3,
// FFMA2 (L1), FADD2, FFMA2 | FSUB2, 2 FSET, 2 LOP3, 2 FADD (L2), 2 FMUL2, 2 FFMA2   (all per pixel PAIR).
template <typename T, int RM, bool ALLV, bool TAIL>
__device__ __forceinline__ void pass_step2(PassAcc<RM>& acc, const T* stage, int pitch, int q0, const bool (&valid)[RM],
                                           const float (&sc)[RM], const float (&tc)[RM], float cntf, float rcnt, bool single) {
    uint64_t xv[RM], iv[RM];
    uint64_t sy = 0ull, si = 0ull;
#pragma unroll
    for (int c = 0; c < RM; ++c) {
        if (ALLV || valid[c]) {
            xv[c] = lds_pair<T>(stage + (size_t)c * pitch + q0);
            const uint64_t y = ffma2(bc2(sc[c]), xv[c], bc2(tc[c]));
            float y0, y1;
            upk2(y, y0, y1);
            iv[c] = pk2(rcp_fast(fmaxf(y0, 1e-3f)), rcp_fast(fmaxf(y1, 1e-3f)));
            sy = fadd2(sy, y);
            si = fadd2(si, iv[c]);
        }
    }
    // target = sum / cnt: reciprocal product + one Newton step (correctly rounded quotient)
    const uint64_t rc = bc2(rcnt), ncnt = bc2(-cntf);
    uint64_t tg = fmul2(sy, rc), td = fmul2(si, rc);
    tg = ffma2(ffma2(ncnt, tg, sy), rc, tg);
    td = ffma2(ffma2(ncnt, td, si), rc, td);
    {
        float g0, g1, d0, d1;
        upk2(tg, g0, g1);
        upk2(td, d0, d1);
        if (TAIL) {
            acc.fa = fadd2(acc.fa, pk2(fabsf(g0), 0.f));
            acc.fb = fadd2(acc.fb, pk2(fabsf(d0), 0.f));
            acc.tmin = fminf(acc.tmin, g0);
            acc.tmax = fmaxf(acc.tmax, g0);
        } else {
            acc.fa = fadd2(acc.fa, pk2(fabsf(g0), fabsf(g1)));
            acc.fb = fadd2(acc.fb, pk2(fabsf(d0), fabsf(d1)));
            acc.tmin = fminf(acc.tmin, fminf(g0, g1));
            acc.tmax = fmaxf(acc.tmax, fmaxf(g0, g1));
        }
    }
    if (single) return;     // y == target exactly: zero data-term gradient (sign(0) = 0) and zero loss
#pragma unroll
    for (int c = 0; c < RM; ++c) {
        if (ALLV || valid[c]) {
            const uint64_t x = xv[c];
            const uint64_t y = ffma2(bc2(sc[c]), x, bc2(tc[c]));
            const uint64_t r1 = fsub2(y, tg);
            const uint64_t sg = signed_mag2(1.0f, TAIL ? 0.0f : 1.0f, r1);        // sign(y - target)
            acc.l1 = ffma2(sg, r1, acc.l1);                                        // |r1|
            acc.A0[c] = fadd2(acc.A0[c], sg);
            acc.A1[c] = ffma2(sg, x, acc.A1[c]);
            const uint64_t in = iv[c];
            const uint64_t r2 = fsub2(in, td);
            float y0, y1, q0f, q1f;
            upk2(y, y0, y1);
            upk2(r2, q0f, q1f);
            acc.l2 += TAIL ? fabsf(q0f) : fabsf(q0f) + fabsf(q1f);
            // d(1/clamp(y, 1e-3))/dy = -1/y^2 where y >= 1e-3, else 0 (the minus sign is applied by the caller)
            const uint64_t sm = signed_mag2(y0 >= 1e-3f ? 1.0f : 0.0f, (!TAIL && y1 >= 1e-3f) ? 1.0f : 0.0f, r2);
            const uint64_t w = fmul2(in, in);
            acc.D0[c] = ffma2(sm, w, acc.D0[c]);
            acc.D1[c] = ffma2(sm, fmul2(w, x), acc.D1[c]);
        }
    }
}

template <typename T, int RM, int kPassThreads, int kCtasPerSm>
__global__ void __launch_bounds__(kPassThreads, kCtasPerSm) coalign_pass_kernel(const PassParams p) {
    constexpr int V = kPassV;
    constexpr int NV = 4 * RM + kFrameStats;
    constexpr int NROUND = (4 * RM + 4 + 31) / 32;       // summed values: 4 per contribution + fa, fb, l1, l2
    extern __shared__ __align__(128) unsigned char pass_smem[];
    __shared__ __align__(8) uint64_t sh_bar;
    __shared__ int sh_row[RM];
    __shared__ float sh_s[RM], sh_t[RM], sh_soft[RM];
    __shared__ float sh_red[kPassThreads / 32][NROUND * 32 + 2];
    __shared__ float sh_tot[NV];
    __shared__ int sh_last;

    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next pass start its x loads early
    const int ch = blockIdx.x;
    const int f = p.f_begin + blockIdx.y;
    const int R = p.R, nt = p.n_total;
    const int px0 = ch * p.px_per_chunk;
    const int npx = min(p.P_pad, px0 + p.px_per_chunk) - px0;     // staged pixels (multiple of 8)
    const int nvalid = min(p.P, px0 + npx) - px0;                 // of which real pixels
    T* stage = reinterpret_cast<T*>(pass_smem);                   // [R][px_per_chunk]

    // ---------------- kick off the bulk copies of this (frame, chunk): nothing below depends on them yet
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pass_smem_u32(&sh_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t total = 0;
        for (int c = 0; c < R; ++c) {
            const int g = c / p.w, j = c - g * p.w;
            const int k = f - j * p.d[g];
            if (k >= 0 && k < p.n[g]) total += (uint32_t)(npx * sizeof(T));
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pass_smem_u32(&sh_bar)), "r"(total) : "memory");
        for (int c = 0; c < R; ++c) {
            const int g = c / p.w, j = c - g * p.w;
            const int k = f - j * p.d[g];
            if (k >= 0 && k < p.n[g]) {
                const T* src = reinterpret_cast<const T*>(p.x) + ((size_t)(p.row0[g] + k) * p.w + j) * p.P_pad + px0;
                bulk_g2s(stage + (size_t)c * p.px_per_chunk, src, (uint32_t)(npx * sizeof(T)), &sh_bar);
            }
        }
    }

    // programmatic dependent launch: everything above touches only x (constant over the iterations) and may run while the
    // previous pass is still finishing; everything below reads or writes what the previous pass wrote
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---------------- prologue: parameters of the <= R snippets touching this frame (overlaps the copies)
    if (tid >= 32 && tid < 32 + RM) {
        const int c = tid - 32;
        int row = -1;
        float s = 0.f, t = 0.f, soft = 0.f;
        if (c < R) {
            const int g = c / p.w, j = c - g * p.w;
            const int d = p.d[g];
            const int k = f - j * d;
            if (k >= 0 && k < p.n[g]) {
                row = p.row0[g] + k;
                s = __ldcg(p.theta_in + row);
                t = __ldcg(p.theta_in + nt + row);
                const float ng = (float)p.n[g];
                if (p.apply_update) {
                    float ms = __ldcg(p.theta_in + 2 * nt + row), mt = __ldcg(p.theta_in + 3 * nt + row);
                    float vs = __ldcg(p.theta_in + 4 * nt + row), vt = __ldcg(p.theta_in + 5 * nt + row);
                    float gs = 0.f, gt = 0.f;
                    for (int jj = 0; jj < p.w; ++jj) {
                        const int fp = k + jj * d;
                        const float2 gg = __ldcg(reinterpret_cast<const float2*>(p.gterm_in) + (size_t)fp * R + g * p.w + jj);
                        gs += gg.x;
                        gt += gg.y;
                    }
                    gs += -2.f * p.lmda2 * fmaxf(1.f - s, 0.f) / ng;
                    gt += 2.f * p.lmda3 * t / ng;
                    adam_update(s, ms, vs, gs, p);
                    adam_update(t, mt, vt, gt, p);
                    if (ch == 0) {
                        // every frame touching this snippet writes the same bits (deterministic recompute)
                        p.theta_out[row] = s;           p.theta_out[nt + row] = t;
                        p.theta_out[2 * nt + row] = ms; p.theta_out[3 * nt + row] = mt;
                        p.theta_out[4 * nt + row] = vs; p.theta_out[5 * nt + row] = vt;
                    }
                }
                if (j == 0) {
                    const float hinge = fmaxf(1.f - s, 0.f);
                    soft = (p.lmda2 * hinge * hinge + p.lmda3 * t * t) / ng;
                }
            }
        }
        sh_row[c] = row;
        sh_s[c] = s;
        sh_t[c] = t;
        sh_soft[c] = soft;
    }
    __syncthreads();
    if (ch == 0 && tid == 0) {
        float soft = 0.f;
        for (int c = 0; c < R; ++c) soft += sh_soft[c];
        p.hist_frame[(size_t)f * 4 + 1] = soft;
    }

    bool valid[RM];
    float sc[RM], tc[RM];
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < RM; ++c) {
        valid[c] = sh_row[c] >= 0;
        sc[c] = sh_s[c];
        tc[c] = sh_t[c];
        cnt += valid[c];
    }
    const float cntf = (float)cnt;
    const float rcnt = __fdiv_rn(1.f, cntf);
    const bool single = cnt <= 1;       // y == target exactly: the data term has zero gradient (sign(0) = 0)

    // ---------------- wait for the staged rows
    {
        uint32_t ok = 0;
        int spins = 0;
        do {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0, 100000;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(ok) : "r"(pass_smem_u32(&sh_bar)) : "memory");
            if (!ok && ++spins > 4000000) __trap();
        } while (!ok);
    }

    // ---------------- main loop: two pixels per thread per step, everything from shared memory.
    // Interior frames have all R contributions (no per-contribution branches); the odd last pixel of a frame
    // is handled by a masked step of one thread.
    PassAcc<RM> acc;
#pragma unroll
    for (int c = 0; c < RM; ++c) { acc.A0[c] = acc.A1[c] = acc.D0[c] = acc.D1[c] = 0ull; }
    acc.fa = acc.fb = acc.l1 = 0ull;
    acc.l2 = 0.f;
    acc.tmin = INFINITY;
    acc.tmax = -INFINITY;
    const int npair = nvalid & ~1;
    if (cnt == R && RM == R) {
        for (int q0 = tid * V; q0 < npair; q0 += kPassThreads * V)
            pass_step2<T, RM, true, false>(acc, stage, p.px_per_chunk, q0, valid, sc, tc, cntf, rcnt, single);
    } else {
        for (int q0 = tid * V; q0 < npair; q0 += kPassThreads * V)
            pass_step2<T, RM, false, false>(acc, stage, p.px_per_chunk, q0, valid, sc, tc, cntf, rcnt, single);
    }
    if ((nvalid & 1) && tid == 0)
        pass_step2<T, RM, false, true>(acc, stage, p.px_per_chunk, nvalid - 1, valid, sc, tc, cntf, rcnt, single);
    auto lanes = [](uint64_t v) { float a, b; upk2(v, a, b); return a + b; };   // even-pixel + odd-pixel partial sums
    float A0[RM], A1[RM], D0[RM], D1[RM];
#pragma unroll
    for (int c = 0; c < RM; ++c) { A0[c] = lanes(acc.A0[c]); A1[c] = lanes(acc.A1[c]); D0[c] = lanes(acc.D0[c]); D1[c] = lanes(acc.D1[c]); }
    const float fa = lanes(acc.fa), fb = lanes(acc.fb), l1 = lanes(acc.l1), l2 = acc.l2, tmin = acc.tmin, tmax = acc.tmax;

    // ---------------- block reduction: butterfly within warps, fixed-order sum across warps
    const int lane = tid & 31, warp = tid >> 5;
    {
        float slot[NROUND][32];
#pragma unroll
        for (int r = 0; r < NROUND; ++r)
#pragma unroll
            for (int i = 0; i < 32; ++i) slot[r][i] = 0.f;
#pragma unroll
        for (int c = 0; c < RM; ++c) {
            slot[(4 * c + 0) / 32][(4 * c + 0) % 32] = A0[c];
            slot[(4 * c + 1) / 32][(4 * c + 1) % 32] = A1[c];
            slot[(4 * c + 2) / 32][(4 * c + 2) % 32] = D0[c];
            slot[(4 * c + 3) / 32][(4 * c + 3) % 32] = D1[c];
        }
        slot[(4 * RM + 0) / 32][(4 * RM + 0) % 32] = fa;
        slot[(4 * RM + 1) / 32][(4 * RM + 1) % 32] = fb;
        slot[(4 * RM + 2) / 32][(4 * RM + 2) % 32] = l1;
        slot[(4 * RM + 3) / 32][(4 * RM + 3) % 32] = l2;
#pragma unroll
        for (int r = 0; r < NROUND; ++r) {
            const float tot = warp_transpose_sum<32>(slot[r], lane);
            // after the butterfly, lane L holds slot index with bits consumed from k = 16 down to 1: index = L
            sh_red[warp][r * 32 + lane] = tot;
        }
        const float mn = warp_min(tmin), mx = warp_max(tmax);
        if (lane == 0) { sh_red[warp][NROUND * 32] = mn; sh_red[warp][NROUND * 32 + 1] = mx; }
    }
    __syncthreads();
    float* part_f = p.part + (size_t)f * p.C * NV;
    if (tid < NV) {
        // NV layout: [4*c + {A0,A1,D0,D1}] for c < RM, then fa, fb, l1, l2, tmin, tmax, pad, pad
        const int q = tid - 4 * RM;
        float v;
        if (q == 4 || q == 5) {
            v = sh_red[0][NROUND * 32 + (q - 4)];
            for (int w = 1; w < kPassThreads / 32; ++w) {
                const float o = sh_red[w][NROUND * 32 + (q - 4)];
                v = (q == 4) ? fminf(v, o) : fmaxf(v, o);
            }
        } else if (q >= 6) {
            v = 0.f;
        } else {
            v = 0.f;
#pragma unroll
            for (int w = 0; w < kPassThreads / 32; ++w) v += sh_red[w][tid];
        }
        if (p.C == 1) sh_tot[tid] = v;
        else part_f[(size_t)ch * NV + tid] = v;
    }
    if (p.C > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned prev = atomicAdd(&p.counter[f], 1u);
            sh_last = (prev == (unsigned)(p.C - 1));
        }
        __syncthreads();
        if (!sh_last) return;
        __threadfence();
        if (tid < NV) {
            float v = __ldcg(part_f + tid);
            const int q = tid - 4 * RM;
            for (int c2 = 1; c2 < p.C; ++c2) {
                const float o = __ldcg(part_f + (size_t)c2 * NV + tid);
                v = (q == 4) ? fminf(v, o) : (q == 5) ? fmaxf(v, o) : v + o;
            }
            sh_tot[tid] = v;
        }
        if (tid == 0) p.counter[f] = 0;
    }
    __syncthreads();

    // ---------------- last chunk of the frame: per-frame scales, gradient terms, loss term ----
    const float a = sh_tot[4 * RM + 0] * p.inv_P;   // mean_p |target|   (depth_aligner.py:197)
    const float b = sh_tot[4 * RM + 1] * p.inv_P;   // mean_p |target_depth|
    const float ia = __fdiv_rn(1.f, a), ib = __fdiv_rn(1.f, b);
    if (tid < R) {
        float2 gg = make_float2(0.f, 0.f);
        if (sh_row[tid] >= 0) {
            const float a0 = sh_tot[4 * tid + 0], a1 = sh_tot[4 * tid + 1];
            const float d0 = sh_tot[4 * tid + 2], d1 = sh_tot[4 * tid + 3];
            gg.x = p.ls_over_Z * (a1 * ia - p.dlw * d1 * ib);   // d loss / d s  (data term)
            gg.y = p.ls_over_Z * (a0 * ia - p.dlw * d0 * ib);   // d loss / d t
        }
        reinterpret_cast<float2*>(p.gterm_out)[(size_t)f * R + tid] = gg;
    }
    if (tid == 0) {
        float* h = p.hist_frame + (size_t)f * 4;
        h[0] = p.ls_over_Z * (sh_tot[4 * RM + 2] * ia + p.dlw * sh_tot[4 * RM + 3] * ib);
        h[2] = sh_tot[4 * RM + 4];
        h[3] = sh_tot[4 * RM + 5];
    }
}

// =============================================================================================
// The fused optimisation pass, contribution-packed organisation (cfg 2 / 3)
// =============================================================================================
// One thread = ONE pixel; the two lanes of every packed fp32x2 value are two CONTRIBUTIONS (i, i + RH) of that pixel
// instead of two pixels of one contribution.  Same packed instruction count per (contribution, pixel), but a thread
// carries accumulators for half as many (contribution, lane) pairs: 4 * RH packed registers instead of 4 * R, which
// removes the spills of the pixel-packed kernel (168 registers, 0.9 M local loads per launch at R = 9) and lets 16
// warps per SM stay resident instead of 12.  Contributions that do not exist for a frame (clip ends) and the odd
// dummy contribution are arithmetic no-ops -- x = 0 from a zero-filled staged row, s = t = 0, a 0 / 1 mask folded into
// the sign constants -- so interior and boundary frames run the same branch-free code.
// The cross-chunk hand-over is a single-thread release / acquire (bar.sync -> fence + ticket by one thread -> bar.sync)
// instead of fences executed by every thread.
template <int RH>
struct PassAccC {
    uint64_t A0[RH], A1[RH], D0[RH], D1[RH];   // lane 0: contribution i, lane 1: contribution i + RH
    uint64_t l1;
    float l2, fa, fb, tmin, tmax;
};

template <typename T>
__device__ __forceinline__ uint64_t lds_two_rows(const T* lo, const T* hi) {
    return pk2(Cvt<T>::to_f(*lo), Cvt<T>::to_f(*hi));
}

template <typename T, int RH, int kPassThreads, int kCtasPerSm>
__global__ void __launch_bounds__(kPassThreads, kCtasPerSm) coalign_pass_cp_kernel(const PassParams p) {
    constexpr int RM = 2 * RH;
    constexpr int NV = 4 * RM + kFrameStats;
    constexpr int NROUND = (4 * RM + 4 + 31) / 32;       // summed values: 4 per contribution + fa, fb, l1, l2
    extern __shared__ __align__(128) unsigned char pass_smem[];
    __shared__ __align__(8) uint64_t sh_bar;
    __shared__ int sh_row[RM];
    __shared__ float sh_s[RM], sh_t[RM], sh_soft[RM];
    __shared__ float sh_red[kPassThreads / 32][NROUND * 32 + 2];
    __shared__ float sh_tot[NV];
    __shared__ int sh_last;

    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next pass start its x loads early
    const int ch = blockIdx.x;
    const int f = p.f_begin + blockIdx.y;
    const int R = p.R, nt = p.n_total;
    const int pitch = p.px_per_chunk;
    const int px0 = ch * pitch;
    const int npx = min(p.P_pad, px0 + pitch) - px0;              // staged pixels (multiple of 8)
    const int nvalid = min(p.P, px0 + npx) - px0;                 // of which real pixels
    T* stage = reinterpret_cast<T*>(pass_smem);                   // [RM][pitch]

    // ---------------- kick off the bulk copies of this (frame, chunk)
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pass_smem_u32(&sh_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t total = 0;
        for (int c = 0; c < R; ++c) {
            const int g = c / p.w, j = c - g * p.w;
            const int k = f - j * p.d[g];
            if (k >= 0 && k < p.n[g]) total += (uint32_t)(npx * sizeof(T));
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pass_smem_u32(&sh_bar)), "r"(total) : "memory");
        for (int c = 0; c < R; ++c) {
            const int g = c / p.w, j = c - g * p.w;
            const int k = f - j * p.d[g];
            if (k >= 0 && k < p.n[g]) {
                const T* src = reinterpret_cast<const T*>(p.x) + ((size_t)(p.row0[g] + k) * p.w + j) * p.P_pad + px0;
                bulk_g2s(stage + (size_t)c * pitch, src, (uint32_t)(npx * sizeof(T)), &sh_bar);
            }
        }
    }

    // programmatic dependent launch: everything above touches only x (constant over the iterations) and may run while the
    // previous pass is still finishing; everything below reads or writes what the previous pass wrote
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---------------- prologue: parameters of the <= R snippets touching this frame (overlaps the copies)
    if (tid >= 32 && tid < 32 + RM) {
        const int c = tid - 32;
        int row = -1;
        float s = 0.f, t = 0.f, soft = 0.f;
        if (c < R) {
            const int g = c / p.w, j = c - g * p.w;
            const int d = p.d[g];
            const int k = f - j * d;
            if (k >= 0 && k < p.n[g]) {
                row = p.row0[g] + k;
                s = __ldcg(p.theta_in + row);
                t = __ldcg(p.theta_in + nt + row);
                const float ng = (float)p.n[g];
                if (p.apply_update) {
                    float ms = __ldcg(p.theta_in + 2 * nt + row), mt = __ldcg(p.theta_in + 3 * nt + row);
                    float vs = __ldcg(p.theta_in + 4 * nt + row), vt = __ldcg(p.theta_in + 5 * nt + row);
                    float gs = 0.f, gt = 0.f;
                    for (int jj = 0; jj < p.w; ++jj) {
                        const int fp = k + jj * d;
                        const float2 gg = __ldcg(reinterpret_cast<const float2*>(p.gterm_in) + (size_t)fp * R + g * p.w + jj);
                        gs += gg.x;
                        gt += gg.y;
                    }
                    gs += -2.f * p.lmda2 * fmaxf(1.f - s, 0.f) / ng;
                    gt += 2.f * p.lmda3 * t / ng;
                    adam_update(s, ms, vs, gs, p);
                    adam_update(t, mt, vt, gt, p);
                    if (ch == 0) {
                        // every frame touching this snippet writes the same bits (deterministic recompute)
                        p.theta_out[row] = s;           p.theta_out[nt + row] = t;
                        p.theta_out[2 * nt + row] = ms; p.theta_out[3 * nt + row] = mt;
                        p.theta_out[4 * nt + row] = vs; p.theta_out[5 * nt + row] = vt;
                    }
                }
                if (j == 0) {
                    const float hinge = fmaxf(1.f - s, 0.f);
                    soft = (p.lmda2 * hinge * hinge + p.lmda3 * t * t) / ng;
                }
            }
        }
        sh_row[c] = row;
        sh_s[c] = row >= 0 ? s : 0.f;       // absent contribution: y = 0 * 0 + 0 (below the clip => no gradient either)
        sh_t[c] = row >= 0 ? t : 0.f;
        sh_soft[c] = soft;
    }
    __syncthreads();
    if (ch == 0 && tid == 0) {
        float soft = 0.f;
        for (int c = 0; c < R; ++c) soft += sh_soft[c];
        p.hist_frame[(size_t)f * 4 + 1] = soft;
    }

    uint64_t sc2[RH], tc2[RH], vm2[RH];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < RH; ++i) {
        const bool v0 = sh_row[i] >= 0, v1 = sh_row[i + RH] >= 0;
        sc2[i] = pk2(sh_s[i], sh_s[i + RH]);
        tc2[i] = pk2(sh_t[i], sh_t[i + RH]);
        vm2[i] = pk2(v0 ? 1.f : 0.f, v1 ? 1.f : 0.f);
        cnt += (int)v0 + (int)v1;
    }
    const float cntf = (float)cnt;
    const float rcnt = __fdiv_rn(1.f, cntf);
    const bool single = cnt <= 1;       // y == target exactly: the data term has zero gradient (sign(0) = 0)
    if (cnt != RM) {                    // CTA-uniform: zero-fill the staged rows of absent contributions (x = 0)
        for (int c = 0; c < RM; ++c) {
            if (sh_row[c] >= 0) continue;
            uint4* z = reinterpret_cast<uint4*>(stage + (size_t)c * pitch);
            for (int q = tid; q < (int)(npx * sizeof(T) / 16); q += kPassThreads) z[q] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
    }

    // ---------------- wait for the staged rows
    {
        uint32_t ok = 0;
        int spins = 0;
        do {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0, 100000;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(ok) : "r"(pass_smem_u32(&sh_bar)) : "memory");
            if (!ok && ++spins > 4000000) __trap();
        } while (!ok);
    }

    // ---------------- main loop: one pixel per thread per step, all contributions, from shared memory
    PassAccC<RH> acc;
#pragma unroll
    for (int i = 0; i < RH; ++i) { acc.A0[i] = acc.A1[i] = acc.D0[i] = acc.D1[i] = 0ull; }
    acc.l1 = 0ull;
    acc.l2 = acc.fa = acc.fb = 0.f;
    acc.tmin = INFINITY;
    acc.tmax = -INFINITY;
    const T* hi_rows = stage + (size_t)RH * pitch;
    for (int q = tid; q < nvalid; q += kPassThreads) {
        uint64_t xv[RH], iv[RH];
        uint64_t sy2 = 0ull, si2 = 0ull;
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            xv[i] = lds_two_rows<T>(stage + (size_t)i * pitch + q, hi_rows + (size_t)i * pitch + q);
            const uint64_t y = ffma2(sc2[i], xv[i], tc2[i]);
            float y0, y1;
            upk2(y, y0, y1);
            iv[i] = pk2(rcp_fast(fmaxf(y0, 1e-3f)), rcp_fast(fmaxf(y1, 1e-3f)));
            sy2 = ffma2(y, vm2[i], sy2);
            si2 = ffma2(iv[i], vm2[i], si2);
        }
        float sya, syb, sia, sib;
        upk2(sy2, sya, syb);
        upk2(si2, sia, sib);
        const float sy = sya + syb, si = sia + sib;
        // target = sum / cnt: reciprocal product + one Newton step (correctly rounded quotient)
        float tg = sy * rcnt, td = si * rcnt;
        tg = fmaf(fmaf(-cntf, tg, sy), rcnt, tg);
        td = fmaf(fmaf(-cntf, td, si), rcnt, td);
        acc.fa += fabsf(tg);
        acc.fb += fabsf(td);
        acc.tmin = fminf(acc.tmin, tg);
        acc.tmax = fmaxf(acc.tmax, tg);
        if (single) continue;           // CTA-uniform
        const uint64_t tg2 = bc2(tg), td2 = bc2(td);
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            const uint64_t x = xv[i];
            const uint64_t y = ffma2(sc2[i], x, tc2[i]);
            const uint64_t r1 = fsub2(y, tg2);
            float v0, v1;
            upk2(vm2[i], v0, v1);
            const uint64_t sg = signed_mag2(v0, v1, r1);                          // sign(y - target), 0 where absent
            acc.l1 = ffma2(sg, r1, acc.l1);                                        // |r1|
            acc.A0[i] = fadd2(acc.A0[i], sg);
            acc.A1[i] = ffma2(sg, x, acc.A1[i]);
            const uint64_t in = iv[i];
            const uint64_t r2 = fsub2(in, td2);
            float y0, y1, q0f, q1f;
            upk2(y, y0, y1);
            upk2(r2, q0f, q1f);
            acc.l2 = fmaf(fabsf(q0f), v0, acc.l2);
            acc.l2 = fmaf(fabsf(q1f), v1, acc.l2);
            // d(1/clamp(y, 1e-3))/dy = -1/y^2 where y >= 1e-3, else 0; absent contributions have y = 0 < 1e-3
            const uint64_t sm = signed_mag2(y0 >= 1e-3f ? 1.0f : 0.0f, y1 >= 1e-3f ? 1.0f : 0.0f, r2);
            const uint64_t w = fmul2(in, in);
            acc.D0[i] = ffma2(sm, w, acc.D0[i]);
            acc.D1[i] = ffma2(sm, fmul2(w, x), acc.D1[i]);
        }
    }

    // ---------------- block reduction: butterfly within warps, fixed-order sum across warps
    const int lane = tid & 31, warp = tid >> 5;
    {
        float slot[NROUND][32];
#pragma unroll
        for (int r = 0; r < NROUND; ++r)
#pragma unroll
            for (int i = 0; i < 32; ++i) slot[r][i] = 0.f;
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            float lo, hi;
            const int c0 = i, c1 = i + RH;
            upk2(acc.A0[i], lo, hi); slot[(4 * c0 + 0) / 32][(4 * c0 + 0) % 32] = lo; slot[(4 * c1 + 0) / 32][(4 * c1 + 0) % 32] = hi;
            upk2(acc.A1[i], lo, hi); slot[(4 * c0 + 1) / 32][(4 * c0 + 1) % 32] = lo; slot[(4 * c1 + 1) / 32][(4 * c1 + 1) % 32] = hi;
            upk2(acc.D0[i], lo, hi); slot[(4 * c0 + 2) / 32][(4 * c0 + 2) % 32] = lo; slot[(4 * c1 + 2) / 32][(4 * c1 + 2) % 32] = hi;
            upk2(acc.D1[i], lo, hi); slot[(4 * c0 + 3) / 32][(4 * c0 + 3) % 32] = lo; slot[(4 * c1 + 3) / 32][(4 * c1 + 3) % 32] = hi;
        }
        float l1a, l1b;
        upk2(acc.l1, l1a, l1b);
        slot[(4 * RM + 0) / 32][(4 * RM + 0) % 32] = acc.fa;
        slot[(4 * RM + 1) / 32][(4 * RM + 1) % 32] = acc.fb;
        slot[(4 * RM + 2) / 32][(4 * RM + 2) % 32] = l1a + l1b;
        slot[(4 * RM + 3) / 32][(4 * RM + 3) % 32] = acc.l2;
#pragma unroll
        for (int r = 0; r < NROUND; ++r) {
            const float tot = warp_transpose_sum<32>(slot[r], lane);
            sh_red[warp][r * 32 + lane] = tot;
        }
        const float mn = warp_min(acc.tmin), mx = warp_max(acc.tmax);
        if (lane == 0) { sh_red[warp][NROUND * 32] = mn; sh_red[warp][NROUND * 32 + 1] = mx; }
    }
    __syncthreads();
    float* part_f = p.part + (size_t)f * p.C * NV;
    if (tid < NV) {
        // NV layout: [4*c + {A0,A1,D0,D1}] for c < RM, then fa, fb, l1, l2, tmin, tmax, pad, pad
        const int q = tid - 4 * RM;
        float v;
        if (q == 4 || q == 5) {
            v = sh_red[0][NROUND * 32 + (q - 4)];
            for (int w = 1; w < kPassThreads / 32; ++w) {
                const float o = sh_red[w][NROUND * 32 + (q - 4)];
                v = (q == 4) ? fminf(v, o) : fmaxf(v, o);
            }
        } else if (q >= 6) {
            v = 0.f;
        } else {
            v = 0.f;
#pragma unroll
            for (int w = 0; w < kPassThreads / 32; ++w) v += sh_red[w][tid];
        }
        if (p.C == 1) sh_tot[tid] = v;
        else part_f[(size_t)ch * NV + tid] = v;
    }
    if (p.C > 1) {
        __syncthreads();                 // the partials of this CTA are written (CTA scope) ...
        if (tid == 0) {
            __threadfence();             // ... one thread publishes them at GPU scope (cumulative release) and takes a ticket
            const unsigned prev = atomicAdd(&p.counter[f], 1u);
            __threadfence();             // acquire side for the last chunk: the other chunks' partials are visible below
            sh_last = (prev == (unsigned)(p.C - 1));
        }
        __syncthreads();
        if (!sh_last) return;
        if (tid < NV) {
            float v = __ldcg(part_f + tid);
            const int q = tid - 4 * RM;
            for (int c2 = 1; c2 < p.C; ++c2) {
                const float o = __ldcg(part_f + (size_t)c2 * NV + tid);
                v = (q == 4) ? fminf(v, o) : (q == 5) ? fmaxf(v, o) : v + o;
            }
            sh_tot[tid] = v;
        }
        if (tid == 0) p.counter[f] = 0;
    }
    __syncthreads();

    // ---------------- last chunk of the frame: per-frame scales, gradient terms, loss term ----
    const float a = sh_tot[4 * RM + 0] * p.inv_P;   // mean_p |target|   (depth_aligner.py:197)
    const float b = sh_tot[4 * RM + 1] * p.inv_P;   // mean_p |target_depth|
    const float ia = __fdiv_rn(1.f, a), ib = __fdiv_rn(1.f, b);
    if (tid < R) {
        float2 gg = make_float2(0.f, 0.f);
        if (sh_row[tid] >= 0) {
            const float a0 = sh_tot[4 * tid + 0], a1 = sh_tot[4 * tid + 1];
            const float d0 = sh_tot[4 * tid + 2], d1 = sh_tot[4 * tid + 3];
            gg.x = p.ls_over_Z * (a1 * ia - p.dlw * d1 * ib);   // d loss / d s  (data term)
            gg.y = p.ls_over_Z * (a0 * ia - p.dlw * d0 * ib);   // d loss / d t
        }
        reinterpret_cast<float2*>(p.gterm_out)[(size_t)f * R + tid] = gg;
    }
    if (tid == 0) {
        float* h = p.hist_frame + (size_t)f * 4;
        h[0] = p.ls_over_Z * (sh_tot[4 * RM + 2] * ia + p.dlw * sh_tot[4 * RM + 3] * ib);
        h[2] = sh_tot[4 * RM + 4];
        h[3] = sh_tot[4 * RM + 5];
    }
}

// Frame-persistent variant of the contribution-packed pass (cfg 5): ONE CTA per frame walks the frame's chunks through a
// two-stage shared-memory ring (the bulk copies of chunk c + 2 are issued as soon as chunk c has been consumed) and keeps
// its accumulators in registers across chunks.  Per frame instead of per (frame, chunk): one parameter prologue, one
// block reduction, no partial sums in global memory, no ticket and no fences.  With ~1400-pixel chunks the per-chunk
// reduction and prologue were a quarter of the instructions of the chunked kernels.
template <typename T, int RH, int kPassThreads, int kCtasPerSm>
__global__ void __launch_bounds__(kPassThreads, kCtasPerSm) coalign_pass_fp_kernel(const PassParams p) {
    constexpr int RM = 2 * RH;
    constexpr int NV = 4 * RM + kFrameStats;
    constexpr int NROUND = (4 * RM + 4 + 31) / 32;       // summed values: 4 per contribution + fa, fb, l1, l2
    extern __shared__ __align__(128) unsigned char pass_smem[];
    __shared__ int sh_row[RM];
    __shared__ float sh_s[RM], sh_t[RM], sh_soft[RM];
    __shared__ float sh_red[kPassThreads / 32][NROUND * 32 + 2];
    __shared__ float sh_tot[NV];

    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next pass start its x loads early
    const int f = p.f_begin + blockIdx.x;
    const int ch = 0;                                             // one CTA owns the whole frame
    const int R = p.R, nt = p.n_total;
    const int pitch = p.px_per_chunk;
    const int nchunk = p.C;
    __shared__ __align__(8) uint64_t sh_full[2];
    __shared__ const T* sh_src[RM];                              // global row start of every contribution (nullptr: absent)
    __shared__ int sh_nrows;
    T* stage0 = reinterpret_cast<T*>(pass_smem);                  // two stages of [RM][pitch]
    const size_t stage_elems = (size_t)RM * pitch;

    // bulk copies of chunk `c` into stage c & 1 (thread 0): the rows of the contributions that exist for this frame
    auto issue_chunk = [&](int c) {
        const int px0 = c * pitch;
        const int npx = min(p.P_pad, px0 + pitch) - px0;          // multiple of 8
        T* st = stage0 + (size_t)(c & 1) * stage_elems;
        uint64_t* bar = &sh_full[c & 1];
        uint32_t total = 0;
        for (int cc = 0; cc < R; ++cc) {
            const int g = cc / p.w, j = cc - g * p.w;
            const int k = f - j * p.d[g];
            if (k >= 0 && k < p.n[g]) total += (uint32_t)(npx * sizeof(T));
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pass_smem_u32(bar)), "r"(total) : "memory");
        for (int cc = 0; cc < R; ++cc) {
            const int g = cc / p.w, j = cc - g * p.w;
            const int k = f - j * p.d[g];
            if (k >= 0 && k < p.n[g]) {
                const T* src = reinterpret_cast<const T*>(p.x) + ((size_t)(p.row0[g] + k) * p.w + j) * p.P_pad + px0;
                bulk_g2s(st + (size_t)cc * pitch, src, (uint32_t)(npx * sizeof(T)), bar);
            }
        }
    };
    // the same from the row pointers the prologue leaves in shared memory (a dozen instructions per row instead of the
    // index arithmetic above: thread 0 is also a compute thread, and its warp is what the per-chunk barrier waits for)
    auto issue_chunk_fast = [&](int c) {
        const int px0 = c * pitch;
        const uint32_t bytes = (uint32_t)((min(p.P_pad, px0 + pitch) - px0) * sizeof(T));
        T* st = stage0 + (size_t)(c & 1) * stage_elems;
        uint64_t* bar = &sh_full[c & 1];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pass_smem_u32(bar)), "r"(bytes * (uint32_t)sh_nrows) : "memory");
#pragma unroll 1
        for (int cc = 0; cc < R; ++cc) {
            const T* src = sh_src[cc];
            if (src != nullptr) bulk_g2s(st + (size_t)cc * pitch, src + px0, bytes, bar);
        }
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pass_smem_u32(&sh_full[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pass_smem_u32(&sh_full[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue_chunk(0);
        if (nchunk > 1) issue_chunk(1);
    }

    // programmatic dependent launch: everything above touches only x (constant over the iterations) and may run while the
    // previous pass is still finishing; everything below reads or writes what the previous pass wrote
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---------------- prologue: parameters of the <= R snippets touching this frame (overlaps the copies)
    if (tid >= 32 && tid < 32 + RM) {
        const int c = tid - 32;
        int row = -1;
        float s = 0.f, t = 0.f, soft = 0.f;
        if (c < R) {
            const int g = c / p.w, j = c - g * p.w;
            const int d = p.d[g];
            const int k = f - j * d;
            if (k >= 0 && k < p.n[g]) {
                row = p.row0[g] + k;
                s = __ldcg(p.theta_in + row);
                t = __ldcg(p.theta_in + nt + row);
                const float ng = (float)p.n[g];
                if (p.apply_update) {
                    float ms = __ldcg(p.theta_in + 2 * nt + row), mt = __ldcg(p.theta_in + 3 * nt + row);
                    float vs = __ldcg(p.theta_in + 4 * nt + row), vt = __ldcg(p.theta_in + 5 * nt + row);
                    float gs = 0.f, gt = 0.f;
                    for (int jj = 0; jj < p.w; ++jj) {
                        const int fp = k + jj * d;
                        const float2 gg = __ldcg(reinterpret_cast<const float2*>(p.gterm_in) + (size_t)fp * R + g * p.w + jj);
                        gs += gg.x;
                        gt += gg.y;
                    }
                    gs += -2.f * p.lmda2 * fmaxf(1.f - s, 0.f) / ng;
                    gt += 2.f * p.lmda3 * t / ng;
                    adam_update(s, ms, vs, gs, p);
                    adam_update(t, mt, vt, gt, p);
                    if (ch == 0) {
                        // every frame touching this snippet writes the same bits (deterministic recompute)
                        p.theta_out[row] = s;           p.theta_out[nt + row] = t;
                        p.theta_out[2 * nt + row] = ms; p.theta_out[3 * nt + row] = mt;
                        p.theta_out[4 * nt + row] = vs; p.theta_out[5 * nt + row] = vt;
                    }
                }
                if (j == 0) {
                    const float hinge = fmaxf(1.f - s, 0.f);
                    soft = (p.lmda2 * hinge * hinge + p.lmda3 * t * t) / ng;
                }
            }
        }
        sh_row[c] = row;
        {
            const int g = c < R ? c / p.w : 0, j = c < R ? c - g * p.w : 0;
            sh_src[c] = row >= 0 ? reinterpret_cast<const T*>(p.x) + ((size_t)row * p.w + j) * p.P_pad : nullptr;
        }
        sh_s[c] = row >= 0 ? s : 0.f;       // absent contribution: y = 0 * 0 + 0 (below the clip => no gradient either)
        sh_t[c] = row >= 0 ? t : 0.f;
        sh_soft[c] = soft;
    }
    __syncthreads();
    if (ch == 0 && tid == 0) {
        float soft = 0.f;
        for (int c = 0; c < R; ++c) soft += sh_soft[c];
        p.hist_frame[(size_t)f * 4 + 1] = soft;
    }

    uint64_t sc2[RH], tc2[RH], vm2[RH];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < RH; ++i) {
        const bool v0 = sh_row[i] >= 0, v1 = sh_row[i + RH] >= 0;
        sc2[i] = pk2(sh_s[i], sh_s[i + RH]);
        tc2[i] = pk2(sh_t[i], sh_t[i + RH]);
        vm2[i] = pk2(v0 ? 1.f : 0.f, v1 ? 1.f : 0.f);
        cnt += (int)v0 + (int)v1;
    }
    if (tid == 0) sh_nrows = cnt;       // read by thread 0 only
    const float cntf = (float)cnt;
    const float rcnt = __fdiv_rn(1.f, cntf);
    const bool single = cnt <= 1;       // y == target exactly: the data term has zero gradient (sign(0) = 0)
    if (cnt != RM) {                    // CTA-uniform: zero-fill the rows of absent contributions in both stages (x = 0)
        for (int c = 0; c < 2 * RM; ++c) {
            if (sh_row[c % RM] >= 0) continue;
            uint4* z = reinterpret_cast<uint4*>(stage0 + (size_t)(c / RM) * stage_elems + (size_t)(c % RM) * pitch);
            for (int q = tid; q < (int)(pitch * sizeof(T) / 16); q += kPassThreads) z[q] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
    }

    PassAccC<RH> acc;
#pragma unroll
    for (int i = 0; i < RH; ++i) { acc.A0[i] = acc.A1[i] = acc.D0[i] = acc.D1[i] = 0ull; }
    acc.l1 = 0ull;
    acc.l2 = acc.fa = acc.fb = 0.f;
    acc.tmin = INFINITY;
    acc.tmax = -INFINITY;
    for (int c = 0; c < nchunk; ++c) {
    {   // wait for chunk c (its stage is used for the (c >> 1)-th time)
        uint32_t ok = 0;
        int spins = 0;
        do {
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2, 100000;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(ok) : "r"(pass_smem_u32(&sh_full[c & 1])), "r"((uint32_t)((c >> 1) & 1)) : "memory");
            if (!ok && ++spins > 4000000) __trap();
        } while (!ok);
    }
    const T* stage = stage0 + (size_t)(c & 1) * stage_elems;
    const T* hi_rows = stage + (size_t)RH * pitch;
    const int px0 = c * pitch;
    const int nvalid = min(p.P, px0 + pitch) - px0;               // real pixels of this chunk
    for (int q = tid; q < nvalid; q += kPassThreads) {
        uint64_t xv[RH], iv[RH];
        uint64_t sy2 = 0ull, si2 = 0ull;
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            xv[i] = lds_two_rows<T>(stage + (size_t)i * pitch + q, hi_rows + (size_t)i * pitch + q);
            const uint64_t y = ffma2(sc2[i], xv[i], tc2[i]);
            float y0, y1;
            upk2(y, y0, y1);
            iv[i] = pk2(rcp_fast(fmaxf(y0, 1e-3f)), rcp_fast(fmaxf(y1, 1e-3f)));
            sy2 = ffma2(y, vm2[i], sy2);
            si2 = ffma2(iv[i], vm2[i], si2);
        }
        float sya, syb, sia, sib;
        upk2(sy2, sya, syb);
        upk2(si2, sia, sib);
        const float sy = sya + syb, si = sia + sib;
        // target = sum / cnt: reciprocal product + one Newton step (correctly rounded quotient)
        float tg = sy * rcnt, td = si * rcnt;
        tg = fmaf(fmaf(-cntf, tg, sy), rcnt, tg);
        td = fmaf(fmaf(-cntf, td, si), rcnt, td);
        acc.fa += fabsf(tg);
        acc.fb += fabsf(td);
        acc.tmin = fminf(acc.tmin, tg);
        acc.tmax = fmaxf(acc.tmax, tg);
        if (single) continue;           // CTA-uniform
        const uint64_t tg2 = bc2(tg), td2 = bc2(td);
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            const uint64_t x = xv[i];
            const uint64_t y = ffma2(sc2[i], x, tc2[i]);
            const uint64_t r1 = fsub2(y, tg2);
            float v0, v1;
            upk2(vm2[i], v0, v1);
            const uint64_t sg = signed_mag2(v0, v1, r1);                          // sign(y - target), 0 where absent
            acc.l1 = ffma2(sg, r1, acc.l1);                                        // |r1|
            acc.A0[i] = fadd2(acc.A0[i], sg);
            acc.A1[i] = ffma2(sg, x, acc.A1[i]);
            const uint64_t in = iv[i];
            const uint64_t r2 = fsub2(in, td2);
            float y0, y1, q0f, q1f;
            upk2(y, y0, y1);
            upk2(r2, q0f, q1f);
            acc.l2 = fmaf(fabsf(q0f), v0, acc.l2);
            acc.l2 = fmaf(fabsf(q1f), v1, acc.l2);
            // d(1/clamp(y, 1e-3))/dy = -1/y^2 where y >= 1e-3, else 0; absent contributions have y = 0 < 1e-3
            const uint64_t sm = signed_mag2(y0 >= 1e-3f ? 1.0f : 0.0f, y1 >= 1e-3f ? 1.0f : 0.0f, r2);
            const uint64_t w = fmul2(in, in);
            acc.D0[i] = ffma2(sm, w, acc.D0[i]);
            acc.D1[i] = ffma2(sm, fmul2(w, x), acc.D1[i]);
        }
    }

    __syncthreads();                                              // every thread is done reading stage c & 1
    if (tid == 0 && c + 2 < nchunk) issue_chunk_fast(c + 2);
    }

    // ---------------- block reduction: butterfly within warps, fixed-order sum across warps
    const int lane = tid & 31, warp = tid >> 5;
    {
        float slot[NROUND][32];
#pragma unroll
        for (int r = 0; r < NROUND; ++r)
#pragma unroll
            for (int i = 0; i < 32; ++i) slot[r][i] = 0.f;
#pragma unroll
        for (int i = 0; i < RH; ++i) {
            float lo, hi;
            const int c0 = i, c1 = i + RH;
            upk2(acc.A0[i], lo, hi); slot[(4 * c0 + 0) / 32][(4 * c0 + 0) % 32] = lo; slot[(4 * c1 + 0) / 32][(4 * c1 + 0) % 32] = hi;
            upk2(acc.A1[i], lo, hi); slot[(4 * c0 + 1) / 32][(4 * c0 + 1) % 32] = lo; slot[(4 * c1 + 1) / 32][(4 * c1 + 1) % 32] = hi;
            upk2(acc.D0[i], lo, hi); slot[(4 * c0 + 2) / 32][(4 * c0 + 2) % 32] = lo; slot[(4 * c1 + 2) / 32][(4 * c1 + 2) % 32] = hi;
            upk2(acc.D1[i], lo, hi); slot[(4 * c0 + 3) / 32][(4 * c0 + 3) % 32] = lo; slot[(4 * c1 + 3) / 32][(4 * c1 + 3) % 32] = hi;
        }
        float l1a, l1b;
        upk2(acc.l1, l1a, l1b);
        slot[(4 * RM + 0) / 32][(4 * RM + 0) % 32] = acc.fa;
        slot[(4 * RM + 1) / 32][(4 * RM + 1) % 32] = acc.fb;
        slot[(4 * RM + 2) / 32][(4 * RM + 2) % 32] = l1a + l1b;
        slot[(4 * RM + 3) / 32][(4 * RM + 3) % 32] = acc.l2;
#pragma unroll
        for (int r = 0; r < NROUND; ++r) {
            const float tot = warp_transpose_sum<32>(slot[r], lane);
            sh_red[warp][r * 32 + lane] = tot;
        }
        const float mn = warp_min(acc.tmin), mx = warp_max(acc.tmax);
        if (lane == 0) { sh_red[warp][NROUND * 32] = mn; sh_red[warp][NROUND * 32 + 1] = mx; }
    }
    __syncthreads();
    if (tid < NV) {
        // NV layout: [4*c + {A0,A1,D0,D1}] for c < RM, then fa, fb, l1, l2, tmin, tmax, pad, pad
        const int q = tid - 4 * RM;
        float v;
        if (q == 4 || q == 5) {
            v = sh_red[0][NROUND * 32 + (q - 4)];
            for (int w = 1; w < kPassThreads / 32; ++w) {
                const float o = sh_red[w][NROUND * 32 + (q - 4)];
                v = (q == 4) ? fminf(v, o) : fmaxf(v, o);
            }
        } else if (q >= 6) {
            v = 0.f;
        } else {
            v = 0.f;
#pragma unroll
            for (int w = 0; w < kPassThreads / 32; ++w) v += sh_red[w][tid];
        }
        sh_tot[tid] = v;
    }
    __syncthreads();

    // ---------------- per-frame scales, gradient terms, loss term ----
    const float a = sh_tot[4 * RM + 0] * p.inv_P;   // mean_p |target|   (depth_aligner.py:197)
    const float b = sh_tot[4 * RM + 1] * p.inv_P;   // mean_p |target_depth|
    const float ia = __fdiv_rn(1.f, a), ib = __fdiv_rn(1.f, b);
    if (tid < R) {
        float2 gg = make_float2(0.f, 0.f);
        if (sh_row[tid] >= 0) {
            const float a0 = sh_tot[4 * tid + 0], a1 = sh_tot[4 * tid + 1];
            const float d0 = sh_tot[4 * tid + 2], d1 = sh_tot[4 * tid + 3];
            gg.x = p.ls_over_Z * (a1 * ia - p.dlw * d1 * ib);   // d loss / d s  (data term)
            gg.y = p.ls_over_Z * (a0 * ia - p.dlw * d0 * ib);   // d loss / d t
        }
        reinterpret_cast<float2*>(p.gterm_out)[(size_t)f * R + tid] = gg;
    }
    if (tid == 0) {
        float* h = p.hist_frame + (size_t)f * 4;
        h[0] = p.ls_over_Z * (sh_tot[4 * RM + 2] * ia + p.dlw * sh_tot[4 * RM + 3] * ib);
        h[2] = sh_tot[4 * RM + 4];
        h[3] = sh_tot[4 * RM + 5];
    }
}


// final Adam update for snippets whose first frame lies in [f_begin, f_end)
__global__ void coalign_final_update_kernel(PassParams p, int f_begin, int f_end, float* s, float* t, float* m_s,
                                            float* m_t, float* v_s, float* v_t) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= p.n_total) return;
    int g = 0;
    for (int q = 1; q < p.G; ++q)
        if (row >= p.row0[q]) g = q;
    const int k = row - p.row0[g];
    const int nt = p.n_total;
    float so = 0.f, to = 0.f, ms = 0.f, mt = 0.f, vs = 0.f, vt = 0.f;
    if (k >= f_begin && k < f_end) {
        so = p.theta_in[row]; to = p.theta_in[nt + row];
        ms = p.theta_in[2 * nt + row]; mt = p.theta_in[3 * nt + row];
        vs = p.theta_in[4 * nt + row]; vt = p.theta_in[5 * nt + row];
        if (p.apply_update) {
            float gs = 0.f, gt = 0.f;
            for (int jj = 0; jj < p.w; ++jj) {
                const int fp = k + jj * p.d[g];
                const float2 gg = reinterpret_cast<const float2*>(p.gterm_in)[(size_t)fp * p.R + g * p.w + jj];
                gs += gg.x;
                gt += gg.y;
            }
            const float ng = (float)p.n[g];
            gs += -2.f * p.lmda2 * fmaxf(1.f - so, 0.f) / ng;
            gt += 2.f * p.lmda3 * to / ng;
            adam_update(so, ms, vs, gs, p);
            adam_update(to, mt, vt, gt, p);
        }
    }
    s[row] = so; t[row] = to; m_s[row] = ms; m_t[row] = mt; v_s[row] = vs; v_t[row] = vt;
}

// history[it] = (sum_f loss_f + soft_f, min_f, max_f) over owned frames; one block per iteration
__global__ void __launch_bounds__(256) coalign_history_kernel(const float* hist_frame, int N, int f_begin, int f_end,
                                                              float* history) {
    const int it = blockIdx.x;
    const float* h = hist_frame + (size_t)it * N * 4;
    float sum = 0.f, mn = INFINITY, mx = -INFINITY;
    for (int f = f_begin + threadIdx.x; f < f_end; f += blockDim.x) {
        sum += h[(size_t)f * 4 + 0] + h[(size_t)f * 4 + 1];
        mn = fminf(mn, h[(size_t)f * 4 + 2]);
        mx = fmaxf(mx, h[(size_t)f * 4 + 3]);
    }
    __shared__ float sh[3][8];
    sum = warp_sum(sum); mn = warp_min(mn); mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = sum; sh[1][threadIdx.x >> 5] = mn; sh[2][threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { sum += sh[0][w]; mn = fminf(mn, sh[1][w]); mx = fmaxf(mx, sh[2][w]); }
        history[3 * it + 0] = sum; history[3 * it + 1] = mn; history[3 * it + 2] = mx;
    }
}

// =============================================================================================
// merge: per-frame mean of ((x - min) * s + t) in the snippet dtype + min/max of the result
// =============================================================================================
struct MergeParams {
    const void* snip[RD_MAX_GROUPS];
    rd_coalign_problem pb;
    const float* s;
    const float* t;
    const float* dev_min;
    void* out;
    RedScratch* sc;
    float* out_minmax;
    int blocks_per_frame;
};

template <typename T>
__global__ void __launch_bounds__(256) merge_kernel(MergeParams mp) {
    const rd_coalign_problem& pb = mp.pb;
    constexpr int VN = Vec16<T>::N;
    __shared__ const T* sh_ptr[RD_MAX_GROUPS * RD_MAX_WINDOW];
    __shared__ float sh_s[RD_MAX_GROUPS * RD_MAX_WINDOW], sh_t[RD_MAX_GROUPS * RD_MAX_WINDOW];
    __shared__ int sh_cnt;
    const int f = blockIdx.y;
    const long long HW = (long long)pb.H * pb.W;
    if (threadIdx.x == 0) {
        int cnt = 0;
        // reference order: for each dilation, boolean mask over [n, w] row-major => k ascending (j descending)
        for (int g = 0; g < pb.G; ++g) {
            for (int j = pb.w - 1; j >= 0; --j) {
                const int k = f - j * pb.d[g];
                if (k >= 0 && k < pb.n[g]) {
                    const int row = pb.row0[g] + k;
                    sh_ptr[cnt] = reinterpret_cast<const T*>(mp.snip[g]) + ((long long)k * pb.w + j) * HW;
                    sh_s[cnt] = Cvt<T>::round(mp.s[row]);   // s.to(dtype)  (depth_aligner.py:248)
                    sh_t[cnt] = Cvt<T>::round(mp.t[row]);   // t.to(dtype)
                    ++cnt;
                }
            }
        }
        sh_cnt = cnt;
    }
    __syncthreads();
    const int cnt = sh_cnt;
    const float cntf = (float)cnt;
    const float mn0 = mp.dev_min[0];
    T* out = reinterpret_cast<T*>(mp.out) + (long long)f * HW;
    float lmin = INFINITY, lmax = -INFINITY;
    // all per-frame planes are 16-byte aligned when HW % VN == 0 and the bases are aligned (checked on host)
    const long long nvec = HW / VN;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
        float acc[VN];
#pragma unroll
        for (int q = 0; q < VN; ++q) acc[q] = 0.f;
        for (int c = 0; c < cnt; ++c) {
            float xv[VN];
            unpack16<T>(ldg_stream16(reinterpret_cast<const uint4*>(sh_ptr[c]) + i), xv);
            const float sT = sh_s[c], tT = sh_t[c];
#pragma unroll
            for (int q = 0; q < VN; ++q) {
                const float a = Cvt<T>::round(__fsub_rn(xv[q], mn0));   // snippet - mn      (:79)
                const float b = Cvt<T>::round(__fmul_rn(a, sT));        // * s.to(dtype)     (:248)
                acc[q] += Cvt<T>::round(__fadd_rn(b, tT));              // + t.to(dtype)     (:249)
            }
        }
        alignas(16) T res[VN];
#pragma unroll
        for (int q = 0; q < VN; ++q) {
            const float m = Cvt<T>::round(__fdiv_rn(acc[q], cntf));  // torch mean: fp32 sum / count, rounded once
            res[q] = Cvt<T>::from_f(m);
            lmin = fminf(lmin, m);
            lmax = fmaxf(lmax, m);
        }
        *reinterpret_cast<uint4*>(out + i * VN) = *reinterpret_cast<const uint4*>(res);
    }
    // scalar tail (HW not a multiple of the vector width)
    for (long long i = nvec * VN + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW;
         i += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < cnt; ++c) {
            const float a = Cvt<T>::round(__fsub_rn(Cvt<T>::to_f(sh_ptr[c][i]), mn0));
            const float b = Cvt<T>::round(__fmul_rn(a, sh_s[c]));
            acc += Cvt<T>::round(__fadd_rn(b, sh_t[c]));
        }
        const float m = Cvt<T>::round(__fdiv_rn(acc, cntf));
        out[i] = Cvt<T>::from_f(m);
        lmin = fminf(lmin, m);
        lmax = fmaxf(lmax, m);
    }
    // min / max of the merged video (for rollingdepth_pipeline.py:316-317)
    __shared__ float shm[2][8];
    __shared__ bool last;
    lmin = warp_min(lmin); lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) { shm[0][threadIdx.x >> 5] = lmin; shm[1][threadIdx.x >> 5] = lmax; }
    __syncthreads();
    const unsigned nblk = gridDim.x * gridDim.y;
    const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lmin = fminf(lmin, shm[0][w]); lmax = fmaxf(lmax, shm[1][w]); }
        mp.sc->part_min[bid] = lmin;
        mp.sc->part_max[bid] = lmax;
        __threadfence();
        last = (atomicAdd(&mp.sc->counter, 1u) == nblk - 1);
    }
    __syncthreads();
    if (last) {
        float a = INFINITY, b = -INFINITY;
        for (unsigned i = threadIdx.x; i < nblk; i += blockDim.x) {
            a = fminf(a, __ldcg(&mp.sc->part_min[i]));
            b = fmaxf(b, __ldcg(&mp.sc->part_max[i]));
        }
        a = warp_min(a); b = warp_max(b);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { shm[0][threadIdx.x >> 5] = a; shm[1][threadIdx.x >> 5] = b; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) { a = fminf(a, shm[0][w]); b = fmaxf(b, shm[1][w]); }
            mp.out_minmax[0] = a;
            mp.out_minmax[1] = b;
            mp.sc->counter = 0;
        }
    }
}

// d -= d.min(); d /= d.max(); d = d * 2 - 1   with a rounding to T after every op
template <typename T>
__global__ void __launch_bounds__(256) renorm_kernel(T* d, long long n, const float* minmax) {
    const float mn = minmax[0];
    const float mx1 = Cvt<T>::round(__fsub_rn(minmax[1], mn));  // max of the shifted tensor (rounding is monotone)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = Cvt<T>::to_f(d[i]);
        v = Cvt<T>::round(__fsub_rn(v, mn));
        v = Cvt<T>::round(__fdiv_rn(v, mx1));
        v = Cvt<T>::round(__fmul_rn(v, 2.0f));
        v = Cvt<T>::round(__fsub_rn(v, 1.0f));
        d[i] = Cvt<T>::from_f(v);
    }
}

// =============================================================================================
// host side
// =============================================================================================
static int check_problem(const rd_coalign_problem* p) {
    RD_CHECK_ARG(p != nullptr, "problem is NULL");
    RD_CHECK_ARG(p->G >= 1 && p->G <= RD_MAX_GROUPS, "G=%d out of range [1,%d]", p->G, RD_MAX_GROUPS);
    RD_CHECK_ARG(p->w >= 1 && p->w <= RD_MAX_WINDOW, "w=%d out of range [1,%d]", p->w, RD_MAX_WINDOW);
    RD_CHECK_ARG(p->N >= 1 && p->H >= 1 && p->W >= 1, "bad N/H/W");
    RD_CHECK_ARG(p->dtype >= RD_F32 && p->dtype <= RD_BF16, "bad dtype %d", p->dtype);
    RD_CHECK_ARG(p->n_total > 0 && p->P > 0 && p->P_pad >= p->P, "problem not initialised (call rd_coalign_problem_init)");
    return RD_OK;
}

struct WsLayout {
    size_t theta, gterm_local, gterm_global, part, counter, hist_frame, flags, total;
    int C, px_per_chunk, RM, NV, cfg;
};

static int pick_rm(int R) {
    const int opts[] = {4, 6, 9, 12, 16};
    for (int o : opts)
        if (R <= o) return o;
    return -1;
}

static int g_pass_cfg_forced = -2;       // -2: not read yet, -1: automatic, 0..6: forced (RD_COALIGN_CFG / rd_coalign_set_config)
static int pass_cfg(const rd_coalign_problem* p) {
    if (g_pass_cfg_forced == -2) {
        const char* e = getenv("RD_COALIGN_CFG");
        g_pass_cfg_forced = e ? atoi(e) : -1;
        if (g_pass_cfg_forced < -1 || g_pass_cfg_forced > 6) g_pass_cfg_forced = -1;
    }
    if (g_pass_cfg_forced >= 0) return g_pass_cfg_forced;
    // measured on B200 (tools/microbench.py align, profiles/r2_coalign_configs.txt):
    //   short clips (about one CTA per SM: the launch is latency-bound, 12.5 us whatever the organisation) -> cfg 0;
    //   long clips, fp32 snippets: frame-persistent contribution-packed kernel (cfg 5: 31.3 / 90.1 us per iteration on
    //   the 77 MB / 208 MB problems against 35.4 / 99.6 for cfg 0); 16-bit snippets: chunked contribution-packed
    //   kernel (cfg 2: 26.9 us against 35.1).
    if (p->N <= 2 * num_sms()) return 0;
    return p->dtype == RD_F32 ? 5 : 2;
}
// cfg 0 / 1: pixel-packed kernel (192 x 2 / 128 x 3 CTAs per SM); cfg 2 / 3: contribution-packed kernel (128 x 4 / 256 x 2);
// cfg 4: contribution-packed, one 512-thread CTA per SM (short clips: about one frame per SM, the launch is latency-bound
// and the only parallelism left is inside the frame)
static int pass_smem_budget(const rd_coalign_problem* p) {
    switch (pass_cfg(p)) {
        case 1: return 65536;
        case 2: return 55 * 1024;
        case 3: return 110 * 1024;
        case 4: return 200 * 1024;
        case 5: return 55 * 1024 / 2;        // cfg 5 / 6: frame-persistent kernel, two stages per CTA (128 x 4 / 256 x 2)
        case 6: return 110 * 1024 / 2;
        default: return 98304;
    }
}
static int pick_rh(int R) {          // contribution pairs of the contribution-packed kernel
    const int opts[] = {2, 3, 4, 5, 6, 8};
    for (int o : opts)
        if ((R + 1) / 2 <= o) return o;
    return -1;
}

// Chunks per frame beyond what shared memory requires.  With one CTA per frame a 150-frame clip is 150 CTAs of 6 warps
// on 148 SMs: every SM runs latency-bound at 1.5 warps per scheduler, and two SMs run two frames.  Splitting frames
// into chunks trades that for the cross-chunk ticket; RD_COALIGN_CHUNKS forces a count (A/B measurements).
static int pass_min_chunks(const rd_coalign_problem* p) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("RD_COALIGN_CHUNKS");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 1) return forced;
    return 1;
}

static WsLayout ws_layout(const rd_coalign_problem* p, int n_iter) {
    WsLayout L{};
    const int R = p->G * p->w;
    L.cfg = pass_cfg(p);
    L.RM = L.cfg >= 2 ? (pick_rh(R) < 0 ? -1 : 2 * pick_rh(R)) : pick_rm(R);
    L.NV = 4 * (L.RM < 0 ? 16 : L.RM) + kFrameStats;
    {   // pixels per (frame, chunk): all rows of a chunk are staged in shared memory at once (cfg 2 / 3 also stage the
        // zero rows of the dummy / absent contributions)
        const int es = p->dtype == RD_F32 ? 4 : 2;
        const int rows = L.cfg >= 2 && L.RM > 0 ? L.RM : R;
        int max_px = pass_smem_budget(p) / (rows * es);
        max_px &= ~7;
        if (max_px < 8) max_px = 8;
        L.C = (p->P_pad + max_px - 1) / max_px;
        if (L.C < 1) L.C = 1;
        const int want = pass_min_chunks(p);          // small problems: more chunks per frame to fill the SMs
        if (want > L.C) L.C = want;
        if (L.C > (p->P_pad + 7) / 8) L.C = (p->P_pad + 7) / 8;
        L.px_per_chunk = (((p->P_pad + L.C - 1) / L.C) + 7) & ~7;
        L.C = (p->P_pad + L.px_per_chunk - 1) / L.px_per_chunk;      // no empty trailing chunk after the rounding
    }
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    L.theta = take(sizeof(float) * 2 * 6 * (size_t)p->n_total);
    L.gterm_local = take(sizeof(float) * 2 * (size_t)p->N * R * 2);
    L.gterm_global = take(sizeof(float) * 2 * (size_t)p->N * R * 2);
    L.part = take(sizeof(float) * (size_t)p->N * L.C * L.NV);
    L.counter = take(sizeof(unsigned) * (size_t)p->N);
    L.hist_frame = take(sizeof(float) * (size_t)(n_iter > 0 ? n_iter : 1) * p->N * 4);
    L.flags = take(256);
    L.total = off;
    return L;
}

struct WsFlags { int single_rank; int n_iter; };

// Launch of one optimisation pass.  `pdl`: programmatic dependent launch -- the pass may begin (bulk loads of x, which no
// pass writes) while the previous pass of the same stream is still running; its griddepcontrol.wait orders everything
// else after the previous pass has completed and flushed.  Only passes that directly follow another pass use it.
template <typename K>
static cudaError_t launch_pass_kernel(K kern, dim3 grid, int threads, size_t smem, cudaStream_t st, bool pdl, const PassParams& pp) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, pp);
}

template <typename T, int RM, int NT, int NCTA>
static int launch_pass_c(const PassParams& pp, dim3 grid, cudaStream_t st) {
    auto kern = coalign_pass_kernel<T, RM, NT, NCTA>;
    const int smem = pp.R * pp.px_per_chunk * (int)sizeof(T);
    static PerDeviceOnce configured;      // cudaFuncSetAttribute is per device and not a stream operation
    if (configured.needed()) {
        RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPassSmemBudgetMax + 1024));
        configured.done();
    }
    RD_CUDA(launch_pass_kernel(kern, grid, NT, (size_t)smem, st, pp.pdl != 0, pp));
    RD_LAUNCH_CHECK("coalign_pass_kernel");
    return RD_OK;
}
template <typename T, int RM>
static int launch_pass_k(const PassParams& pp, int cfg, dim3 grid, cudaStream_t st) {
    if (cfg == 1) return launch_pass_c<T, RM, 128, 3>(pp, grid, st);
    return launch_pass_c<T, RM, 192, 2>(pp, grid, st);
}

template <typename T, int RH, int NT, int NCTA>
static int launch_pass_cp(const PassParams& pp, dim3 grid, cudaStream_t st) {
    auto kern = coalign_pass_cp_kernel<T, RH, NT, NCTA>;
    const int smem = 2 * RH * pp.px_per_chunk * (int)sizeof(T);
    static PerDeviceOnce configured;      // cudaFuncSetAttribute is per device and not a stream operation
    if (configured.needed()) {
        RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, NCTA == 1 ? 204 * 1024 : 112 * 1024));
        configured.done();
    }
    RD_CUDA(launch_pass_kernel(kern, grid, NT, (size_t)smem, st, pp.pdl != 0, pp));
    RD_LAUNCH_CHECK("coalign_pass_cp_kernel");
    return RD_OK;
}
template <typename T, int RH, int NT, int NCTA>
static int launch_pass_fp(const PassParams& pp, int n_frames, cudaStream_t st) {
    auto kern = coalign_pass_fp_kernel<T, RH, NT, NCTA>;
    const int smem = 2 * 2 * RH * pp.px_per_chunk * (int)sizeof(T);
    static PerDeviceOnce configured;      // cudaFuncSetAttribute is per device and not a stream operation
    if (configured.needed()) {
        RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        configured.done();
    }
    RD_CUDA(launch_pass_kernel(kern, dim3((unsigned)n_frames), NT, (size_t)smem, st, pp.pdl != 0, pp));
    RD_LAUNCH_CHECK("coalign_pass_fp_kernel");
    return RD_OK;
}
template <typename T, int RH>
static int launch_pass_cpk(const PassParams& pp, int cfg, dim3 grid, cudaStream_t st) {
    if (cfg == 6) return launch_pass_fp<T, RH, 256, 2>(pp, (int)grid.y, st);
    if (cfg == 5) return launch_pass_fp<T, RH, 128, 4>(pp, (int)grid.y, st);
    if (cfg == 4) return launch_pass_cp<T, RH, 512, 1>(pp, grid, st);
    if (cfg == 3) return launch_pass_cp<T, RH, 256, 2>(pp, grid, st);
    return launch_pass_cp<T, RH, 128, 4>(pp, grid, st);
}

template <typename T>
static int launch_pass_t(const PassParams& pp, int RM, int cfg, dim3 grid, cudaStream_t st) {
    if (cfg >= 2) {
        switch (RM / 2) {
            case 2: return launch_pass_cpk<T, 2>(pp, cfg, grid, st);
            case 3: return launch_pass_cpk<T, 3>(pp, cfg, grid, st);
            case 4: return launch_pass_cpk<T, 4>(pp, cfg, grid, st);
            case 5: return launch_pass_cpk<T, 5>(pp, cfg, grid, st);
            case 6: return launch_pass_cpk<T, 6>(pp, cfg, grid, st);
            case 8: return launch_pass_cpk<T, 8>(pp, cfg, grid, st);
            default: set_error("unsupported G*w"); return RD_ERR_UNSUPPORTED;
        }
    }
    switch (RM) {
        case 4:  return launch_pass_k<T, 4>(pp, cfg, grid, st);
        case 6:  return launch_pass_k<T, 6>(pp, cfg, grid, st);
        case 9:  return launch_pass_k<T, 9>(pp, cfg, grid, st);
        case 12: return launch_pass_k<T, 12>(pp, cfg, grid, st);
        case 16: return launch_pass_k<T, 16>(pp, cfg, grid, st);
        default: set_error("unsupported G*w"); return RD_ERR_UNSUPPORTED;
    }
}

static void fill_pass_params(PassParams& pp, const rd_coalign_problem* p, const rd_coalign_hparams* hp, const WsLayout& L,
                             char* ws, int it, int step, bool single) {
    const int R = p->G * p->w;
    pp.G = p->G; pp.w = p->w; pp.N = p->N; pp.P = p->P; pp.P_pad = p->P_pad; pp.R = R; pp.C = L.C;
    pp.n_total = p->n_total; pp.px_per_chunk = L.px_per_chunk;
    for (int g = 0; g < RD_MAX_GROUPS; ++g) { pp.d[g] = p->d[g]; pp.n[g] = p->n[g]; pp.row0[g] = p->row0[g]; }
    pp.apply_update = it > 0;
    // torch.optim.Adam: python-double bias corrections, then fp32 tensor math
    const double b1 = hp->beta1, b2 = hp->beta2;
    const double k = (double)step;  // the update applied here is update number `step`
    pp.step_size = (float)((double)hp->lr / (1.0 - pow(b1, k > 0 ? k : 1.0)));
    pp.bc2_sqrt = (float)sqrt(1.0 - pow(b2, k > 0 ? k : 1.0));
    pp.beta1 = hp->beta1; pp.beta2 = hp->beta2; pp.eps = hp->eps;
    pp.lmda2 = hp->lmda2; pp.lmda3 = hp->lmda3; pp.dlw = hp->depth_loss_weight;
    const double Z = (double)R * (double)p->N * (double)p->P;
    pp.ls_over_Z = (float)((double)hp->loss_scale / Z);
    pp.inv_P = (float)(1.0 / (double)p->P);
    float* theta = reinterpret_cast<float*>(ws + L.theta);
    const size_t th = 6 * (size_t)p->n_total;
    // theta[(it-1)&1] holds theta_{it-1}; pass it writes theta_it to theta[it&1]; pass 0 reads theta[0]
    pp.theta_in = theta + (it > 0 ? ((it - 1) & 1) : 0) * th;
    pp.theta_out = theta + (it & 1) * th;
    const size_t gt = (size_t)p->N * R * 2;
    float* gl = reinterpret_cast<float*>(ws + L.gterm_local);
    float* gg = single ? gl : reinterpret_cast<float*>(ws + L.gterm_global);
    pp.gterm_in = gg + ((it > 0 ? it - 1 : 0) & 1) * gt;
    pp.gterm_out = gl + (it & 1) * gt;
    pp.part = reinterpret_cast<float*>(ws + L.part);
    pp.counter = reinterpret_cast<unsigned*>(ws + L.counter);
    pp.hist_frame = reinterpret_cast<float*>(ws + L.hist_frame) + (size_t)it * p->N * 4;
}

}  // namespace rd

using namespace rd;

extern "C" {

int rd_coalign_problem_init(rd_coalign_problem* p) {
    RD_CHECK_ARG(p != nullptr, "problem is NULL");
    RD_CHECK_ARG(p->G >= 1 && p->G <= RD_MAX_GROUPS, "G=%d out of range [1,%d]", p->G, RD_MAX_GROUPS);
    RD_CHECK_ARG(p->w >= 1 && p->w <= RD_MAX_WINDOW, "w=%d out of range [1,%d]", p->w, RD_MAX_WINDOW);
    RD_CHECK_ARG(p->N >= 1 && p->H >= 1 && p->W >= 1, "bad N/H/W (%d,%d,%d)", p->N, p->H, p->W);
    RD_CHECK_ARG(p->factor >= 1, "factor must be >= 1");
    // the reference slices [border:-border]; border == 0 would give an empty tensor there
    RD_CHECK_ARG(p->border >= 1, "border must be >= 1 (the reference's [border:-border] slice is empty for 0)");
    RD_CHECK_ARG(p->H > 2 * p->border && p->W > 2 * p->border, "frame smaller than the border crop");
    int row = 0;
    for (int g = 0; g < p->G; ++g) {
        RD_CHECK_ARG(p->d[g] >= 1, "dilation[%d]=%d must be >= 1", g, p->d[g]);
        p->n[g] = p->N - (p->w - 1) * p->d[g];
        RD_CHECK_ARG(p->n[g] >= 1, "dilation %d too large for %d frames with window %d", p->d[g], p->N, p->w);
        p->row0[g] = row;
        row += p->n[g];
    }
    for (int g = p->G; g < RD_MAX_GROUPS; ++g) { p->d[g] = 1; p->n[g] = 0; p->row0[g] = row; }
    p->n_total = row;
    p->Hs = (p->H - 2 * p->border + p->factor - 1) / p->factor;
    p->Ws = (p->W - 2 * p->border + p->factor - 1) / p->factor;
    p->P = p->Hs * p->Ws;
    p->P_pad = (p->P + 7) & ~7;
    return RD_OK;
}

int64_t rd_reduce_scratch_bytes(void) { return (int64_t)sizeof(RedScratch); }

int rd_global_min(const void* const* host_snip_ptrs, const rd_coalign_problem* p, float* out_min, void* scratch,
                  void* stream) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(host_snip_ptrs && out_min && scratch, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SegList segs{};
    segs.nseg = p->G;
    long long total = 0;
    const size_t es = p->dtype == RD_F32 ? 4 : 2;
    for (int g = 0; g < p->G; ++g) {
        RD_CHECK_ARG(host_snip_ptrs[g] != nullptr, "snippet pointer %d is NULL", g);
        RD_CHECK_ARG((reinterpret_cast<uintptr_t>(host_snip_ptrs[g]) % es) == 0, "snippet pointer %d misaligned", g);
        segs.ptr[g] = host_snip_ptrs[g];
        segs.count[g] = (long long)p->n[g] * p->w * p->H * p->W;
        total += segs.count[g];
    }
    RedScratch* sc = reinterpret_cast<RedScratch*>(scratch);
    RD_CUDA(cudaMemsetAsync(&sc->counter, 0, sizeof(unsigned), st));
    long long want = (total / (16 / (long long)es) + 256 * 8 - 1) / (256 * 8);
    int blocks = (int)(want < 1 ? 1 : (want > num_sms() * 8 ? num_sms() * 8 : want));
    if (blocks > kRedMaxBlocks) blocks = kRedMaxBlocks;
    switch (p->dtype) {
        case RD_F32: global_min_kernel<float><<<blocks, 256, 0, st>>>(segs, sc, out_min); break;
        case RD_F16: global_min_kernel<__half><<<blocks, 256, 0, st>>>(segs, sc, out_min); break;
        default:     global_min_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(segs, sc, out_min); break;
    }
    RD_LAUNCH_CHECK("global_min_kernel");
    return RD_OK;
}

int rd_coalign_prep(const void* const* host_snip_ptrs, const rd_coalign_problem* p, const float* dev_min,
                    void* x_small, void* stream) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(host_snip_ptrs && dev_min && x_small, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    PrepParams pp{};
    pp.pb = *p;
    for (int g = 0; g < p->G; ++g) {
        RD_CHECK_ARG(host_snip_ptrs[g] != nullptr, "snippet pointer %d is NULL", g);
        pp.snip[g] = host_snip_ptrs[g];
    }
    const long long total = (long long)p->n_total * p->w * p->P_pad;
    long long want = (total + 255) / 256;
    const int blocks = (int)(want > num_sms() * 16 ? num_sms() * 16 : want);
    switch (p->dtype) {
        case RD_F32: prep_kernel<float><<<blocks, 256, 0, st>>>(pp, dev_min, (float*)x_small); break;
        case RD_F16: prep_kernel<__half><<<blocks, 256, 0, st>>>(pp, dev_min, (__half*)x_small); break;
        default:     prep_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(pp, dev_min, (__nv_bfloat16*)x_small); break;
    }
    RD_LAUNCH_CHECK("prep_kernel");
    return RD_OK;
}

int rd_coalign_set_config(int cfg) {
    RD_CHECK_ARG(cfg >= -1 && cfg <= 6, "cfg must be -1 (automatic) or 0..6");
    (void)pass_cfg;                       // make sure the environment default has been read first
    if (g_pass_cfg_forced == -2) { rd_coalign_problem dummy{}; dummy.n_total = 1; dummy.w = 1; dummy.P_pad = 8; (void)pass_cfg(&dummy); }
    const int prev = g_pass_cfg_forced;
    g_pass_cfg_forced = cfg;
    return prev + 2;                      // >= 1 on success (previous setting + 2), so that negative values stay errors
}

int64_t rd_coalign_workspace_bytes(const rd_coalign_problem* p, int n_iter) {
    if (check_problem(p) != RD_OK) return RD_ERR_BAD_ARG;
    return (int64_t)ws_layout(p, n_iter).total;
}

int rd_coalign_begin(const rd_coalign_problem* p, const float* s, const float* t, const float* m_s, const float* m_t,
                     const float* v_s, const float* v_t, int n_iter, int single_rank, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(s && t && m_s && m_t && v_s && v_t && workspace, "NULL pointer");
    RD_CHECK_ARG(n_iter >= 0, "n_iter < 0");
    const WsLayout L = ws_layout(p, n_iter);
    if (L.RM < 0) { set_error("G*w = %d > 16 is not instantiated", p->G * p->w); return RD_ERR_UNSUPPORTED; }
    if ((int64_t)L.total > workspace_bytes) {
        set_error("workspace too small: need %lld bytes, got %lld", (long long)L.total, (long long)workspace_bytes);
        return RD_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = reinterpret_cast<char*>(workspace);
    float* theta = reinterpret_cast<float*>(ws + L.theta);
    const size_t nt = (size_t)p->n_total, b = nt * sizeof(float);
    const float* src[6] = {s, t, m_s, m_t, v_s, v_t};
    for (int i = 0; i < 6; ++i) RD_CUDA(cudaMemcpyAsync(theta + i * nt, src[i], b, cudaMemcpyDeviceToDevice, st));
    RD_CUDA(cudaMemsetAsync(ws + L.gterm_local, 0, L.part - L.gterm_local, st));   // both gterm buffers
    RD_CUDA(cudaMemsetAsync(ws + L.counter, 0, sizeof(unsigned) * (size_t)p->N, st));
    RD_CUDA(cudaMemsetAsync(ws + L.hist_frame, 0, sizeof(float) * (size_t)(n_iter > 0 ? n_iter : 1) * p->N * 4, st));
    WsFlags fl{single_rank, n_iter};
    RD_CUDA(cudaMemcpyAsync(ws + L.flags, &fl, sizeof(fl), cudaMemcpyHostToDevice, st));
    return RD_OK;
}

static int pass_impl(const void* x_small, const rd_coalign_problem* p, const rd_coalign_hparams* hp, int it, int step0,
                     int f_begin, int f_end, void* workspace, void* stream, bool single, int n_iter_layout, bool pdl = false) {
    const WsLayout L = ws_layout(p, n_iter_layout);
    PassParams pp{};
    pp.x = x_small;
    pp.f_begin = f_begin;
    pp.pdl = pdl ? 1 : 0;
    fill_pass_params(pp, p, hp, L, reinterpret_cast<char*>(workspace), it, step0 + it, single);
    dim3 grid(L.C, f_end - f_begin);
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->dtype) {
        case RD_F32: return launch_pass_t<float>(pp, L.RM, L.cfg, grid, st);
        case RD_F16: return launch_pass_t<__half>(pp, L.RM, L.cfg, grid, st);
        default:     return launch_pass_t<__nv_bfloat16>(pp, L.RM, L.cfg, grid, st);
    }
}

// n_iter used for the layout is irrelevant for every offset except the (last) flags block, which the
// pass does not touch, so the pass entry point does not need it.
int rd_coalign_pass(const void* x_small, const rd_coalign_problem* p, const rd_coalign_hparams* hp, int it, int step0,
                    int f_begin, int f_end, void* workspace, void* stream) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(x_small && hp && workspace, "NULL pointer");
    RD_CHECK_ARG(it >= 0 && 0 <= f_begin && f_begin < f_end && f_end <= p->N, "bad it / frame range [%d,%d)", f_begin, f_end);
    const bool single = (f_begin == 0 && f_end == p->N);
    return pass_impl(x_small, p, hp, it, step0, f_begin, f_end, workspace, stream, single, 1);
}

int rd_coalign_exchange(const rd_coalign_problem* p, void* workspace, int it, float** send, float** recv,
                        int64_t* n_floats) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(workspace && send && recv && n_floats, "NULL pointer");
    const WsLayout L = ws_layout(p, 1);
    const size_t gt = (size_t)p->N * p->G * p->w * 2;
    char* ws = reinterpret_cast<char*>(workspace);
    *send = reinterpret_cast<float*>(ws + L.gterm_local) + (it & 1) * gt;
    *recv = reinterpret_cast<float*>(ws + L.gterm_global) + (it & 1) * gt;
    *n_floats = (int64_t)gt;
    return RD_OK;
}

int rd_coalign_end(const rd_coalign_problem* p, const rd_coalign_hparams* hp, int n_iter, int step0, int f_begin,
                   int f_end, float* s, float* t, float* m_s, float* m_t, float* v_s, float* v_t, float* history,
                   void* workspace, void* stream) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(hp && s && t && m_s && m_t && v_s && v_t && workspace, "NULL pointer");
    RD_CHECK_ARG(n_iter >= 0 && 0 <= f_begin && f_begin < f_end && f_end <= p->N, "bad n_iter / frame range");
    const WsLayout L = ws_layout(p, n_iter);
    cudaStream_t st = (cudaStream_t)stream;
    const bool single = (f_begin == 0 && f_end == p->N);
    PassParams pp{};
    // "pass n_iter" would apply update number step0 + n_iter from the terms of pass n_iter - 1
    fill_pass_params(pp, p, hp, L, reinterpret_cast<char*>(workspace), n_iter, step0 + n_iter, single);
    coalign_final_update_kernel<<<(p->n_total + 127) / 128, 128, 0, st>>>(pp, f_begin, f_end, s, t, m_s, m_t, v_s, v_t);
    RD_LAUNCH_CHECK("coalign_final_update_kernel");
    if (n_iter > 0 && history != nullptr) {
        coalign_history_kernel<<<n_iter, 256, 0, st>>>(reinterpret_cast<const float*>((char*)workspace + L.hist_frame),
                                                       p->N, f_begin, f_end, history);
        RD_LAUNCH_CHECK("coalign_history_kernel");
    }
    return RD_OK;
}

int rd_coalign_optimize(const void* x_small, const rd_coalign_problem* p, const rd_coalign_hparams* hp, float* s,
                        float* t, float* m_s, float* m_t, float* v_s, float* v_t, int step0, int n_iter,
                        float* history, void* workspace, int64_t workspace_bytes, void* stream) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(x_small && hp, "NULL pointer");
    if (int e = rd_coalign_begin(p, s, t, m_s, m_t, v_s, v_t, n_iter, 1, workspace, workspace_bytes, stream)) return e;
    static int pdl_env = -1;             // RD_COALIGN_PDL=0 disables the programmatic dependent launches (A/B)
    if (pdl_env < 0) { const char* e = getenv("RD_COALIGN_PDL"); pdl_env = e ? (atoi(e) != 0) : 1; }
    for (int it = 0; it < n_iter; ++it) {
        // pass `it` directly follows pass `it - 1` in the stream: it may start loading x while that one drains
        if (int e = pass_impl(x_small, p, hp, it, step0, 0, p->N, workspace, stream, true, n_iter, pdl_env && it > 0)) return e;
    }
    return rd_coalign_end(p, hp, n_iter, step0, 0, p->N, s, t, m_s, m_t, v_s, v_t, history, workspace, stream);
}

int rd_merge(const void* const* host_snip_ptrs, const rd_coalign_problem* p, const float* dev_min, const float* s,
             const float* t, void* merged, float* out_minmax, void* scratch, void* stream) {
    if (int e = check_problem(p)) return e;
    RD_CHECK_ARG(host_snip_ptrs && dev_min && s && t && merged && out_minmax && scratch, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    MergeParams mp{};
    mp.pb = *p;
    const size_t es = p->dtype == RD_F32 ? 4 : 2;
    const long long HW = (long long)p->H * p->W;
    for (int g = 0; g < p->G; ++g) {
        RD_CHECK_ARG(host_snip_ptrs[g] != nullptr, "snippet pointer %d is NULL", g);
        RD_CHECK_ARG((reinterpret_cast<uintptr_t>(host_snip_ptrs[g]) & 15) == 0, "snippet pointer %d must be 16-byte aligned", g);
        mp.snip[g] = host_snip_ptrs[g];
    }
    RD_CHECK_ARG((reinterpret_cast<uintptr_t>(merged) & 15) == 0, "merged must be 16-byte aligned");
    RD_CHECK_ARG((HW * (long long)es) % 16 == 0, "H*W*sizeof must be a multiple of 16 bytes (got H=%d W=%d)", p->H, p->W);
    mp.s = s; mp.t = t; mp.dev_min = dev_min; mp.out = merged;
    mp.sc = reinterpret_cast<RedScratch*>(scratch);
    mp.out_minmax = out_minmax;
    RD_CUDA(cudaMemsetAsync(&mp.sc->counter, 0, sizeof(unsigned), st));
    const long long nvec = HW / (16 / (long long)es);
    int bpf = (int)((nvec + 256 * 2 - 1) / (256 * 2));   // ~2 vectors per thread
    if (bpf < 1) bpf = 1;
    RD_CHECK_ARG(p->N <= kRedMaxBlocks, "too many frames for one merge launch (N=%d > %d)", p->N, kRedMaxBlocks);
    while ((long long)bpf * p->N > kRedMaxBlocks && bpf > 1) bpf = (bpf + 1) / 2;
    if ((long long)bpf * p->N > kRedMaxBlocks) bpf = 1;
    mp.blocks_per_frame = bpf;
    dim3 grid(bpf, p->N);
    switch (p->dtype) {
        case RD_F32: merge_kernel<float><<<grid, 256, 0, st>>>(mp); break;
        case RD_F16: merge_kernel<__half><<<grid, 256, 0, st>>>(mp); break;
        default:     merge_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(mp); break;
    }
    RD_LAUNCH_CHECK("merge_kernel");
    return RD_OK;
}

int rd_renorm(void* depth, int64_t n, const float* dev_minmax, rd_dtype dtype, void* stream) {
    RD_CHECK_ARG(depth && dev_minmax && n > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    long long want = (n + 256 * 4 - 1) / (256 * 4);
    const int blocks = (int)(want > num_sms() * 16 ? num_sms() * 16 : want);
    switch (dtype) {
        case RD_F32: renorm_kernel<float><<<blocks, 256, 0, st>>>((float*)depth, n, dev_minmax); break;
        case RD_F16: renorm_kernel<__half><<<blocks, 256, 0, st>>>((__half*)depth, n, dev_minmax); break;
        case RD_BF16: renorm_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)depth, n, dev_minmax); break;
        default: set_error("bad dtype"); return RD_ERR_BAD_ARG;
    }
    RD_LAUNCH_CHECK("renorm_kernel");
    return RD_OK;
}

}  // extern "C"
