// fp32 cross-frame attention (the `paper` preset and the CPU-comparable config run in fp32).
// True fp32 FMA arithmetic -- single-pass TF32/BF16 tensor-core MMA is not accurate enough for the
// 1e-4 max-abs bound against the reference's fp32 SDPA path (attention_processor.py:2251-2253).
// One thread per query row (q and the output accumulator live in registers), keys/values staged
// through shared memory in tiles of 32 and read as warp-wide broadcasts.
#include <math.h>

#include "common.cuh"

namespace rd {

constexpr int kF32Threads = 128;
constexpr int kF32BK = 32;
constexpr int kF32D = 64;

struct F32Params {
    const float *q, *k, *v;
    float* o;
    long long qs[3], ks[3], vs[3], os[3];   // batch, seq, head strides (elements)
    int Sq, Sk;
    float scale_log2;
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kF32Threads, 2) xattn_f32_kernel(const F32Params p) {
    __shared__ __align__(16) float sK[kF32BK][kF32D];
    __shared__ __align__(16) float sV[kF32BK][kF32D];
    const int tid = threadIdx.x;
    const int head = blockIdx.y, batch = blockIdx.z;
    const int row = blockIdx.x * kF32Threads + tid;
    const bool live = row < p.Sq;
    const float* qp = p.q + batch * p.qs[0] + (long long)(live ? row : 0) * p.qs[1] + head * p.qs[2];
    const float* kbase = p.k + batch * p.ks[0] + head * p.ks[2];
    const float* vbase = p.v + batch * p.vs[0] + head * p.vs[2];

    float qr[kF32D], acc[kF32D];
#pragma unroll
    for (int d = 0; d < kF32D; d += 4) {
        const float4 t = *reinterpret_cast<const float4*>(qp + d);
        qr[d] = t.x; qr[d + 1] = t.y; qr[d + 2] = t.z; qr[d + 3] = t.w;
        acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    const float sl2 = p.scale_log2;

    for (int k0 = 0; k0 < p.Sk; k0 += kF32BK) {
        // cooperative, coalesced tile load (zero-fill past Sk)
#pragma unroll
        for (int it = 0; it < (kF32BK * kF32D / 4) / kF32Threads; ++it) {
            const int idx = it * kF32Threads + tid;      // float4 index
            const int kr = idx / (kF32D / 4), c4 = idx % (kF32D / 4);
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + kr < p.Sk) {
                kv = *reinterpret_cast<const float4*>(kbase + (long long)(k0 + kr) * p.ks[1] + c4 * 4);
                vv = *reinterpret_cast<const float4*>(vbase + (long long)(k0 + kr) * p.vs[1] + c4 * 4);
            }
            *reinterpret_cast<float4*>(&sK[kr][c4 * 4]) = kv;
            *reinterpret_cast<float4*>(&sV[kr][c4 * 4]) = vv;
        }
        __syncthreads();

        float s[kF32BK];
        float mt = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < kF32BK; ++kk) {
            float dot = 0.f;
#pragma unroll
            for (int d = 0; d < kF32D; d += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&sK[kk][d]);   // broadcast
                dot = fmaf(qr[d], t.x, dot);
                dot = fmaf(qr[d + 1], t.y, dot);
                dot = fmaf(qr[d + 2], t.z, dot);
                dot = fmaf(qr[d + 3], t.w, dot);
            }
            s[kk] = (k0 + kk < p.Sk) ? dot : -INFINITY;
            mt = fmaxf(mt, s[kk]);
        }
        const float m_new = fmaxf(m, mt);
        const float f = ex2f((m - m_new) * sl2);   // m = -inf on the first tile -> f = 0
        l *= f;
#pragma unroll
        for (int d = 0; d < kF32D; ++d) acc[d] *= f;
        const float mb = m_new * sl2;
#pragma unroll
        for (int kk = 0; kk < kF32BK; ++kk) {
            const float pr = ex2f(fmaf(s[kk], sl2, -mb));
            l += pr;
#pragma unroll
            for (int d = 0; d < kF32D; d += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&sV[kk][d]);
                acc[d] = fmaf(pr, t.x, acc[d]);
                acc[d + 1] = fmaf(pr, t.y, acc[d + 1]);
                acc[d + 2] = fmaf(pr, t.z, acc[d + 2]);
                acc[d + 3] = fmaf(pr, t.w, acc[d + 3]);
            }
        }
        m = m_new;
        __syncthreads();
    }
    if (live) {
        float* op = p.o + batch * p.os[0] + (long long)row * p.os[1] + head * p.os[2];
        const float inv = 1.0f / l;
#pragma unroll
        for (int d = 0; d < kF32D; d += 4)
            *reinterpret_cast<float4*>(op + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
    }
}

int xattn_f32_launch(const float* q, const float* k, const float* v, float* o, int B, int H, int Sq, int Sk,
                     const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, float scale,
                     cudaStream_t st) {
    for (int i = 0; i < 3; ++i)
        RD_CHECK_ARG(qs[i] % 4 == 0 && ks[i] % 4 == 0 && vs[i] % 4 == 0 && os[i] % 4 == 0,
                     "fp32 strides must be multiples of 4 elements");
    RD_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)o % 16) == 0,
                 "q/k/v/o must be 16-byte aligned");
    F32Params p{};
    p.q = q; p.k = k; p.v = v; p.o = o;
    for (int i = 0; i < 3; ++i) { p.qs[i] = qs[i]; p.ks[i] = ks[i]; p.vs[i] = vs[i]; p.os[i] = os[i]; }
    p.Sq = Sq; p.Sk = Sk;
    p.scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((Sq + kF32Threads - 1) / kF32Threads, H, B);
    xattn_f32_kernel<<<grid, kF32Threads, 0, st>>>(p);
    RD_LAUNCH_CHECK("xattn_f32_kernel");
    return RD_OK;
}

}  // namespace rd
