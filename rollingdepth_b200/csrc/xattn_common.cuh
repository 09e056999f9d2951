// Helpers shared by the two tcgen05 cross-frame attention kernels (xattn_sm100.cu: two query tiles per CTA,
// xattn_sm100_q3.cu: three): packed fp32x2 arithmetic, MUFU / polynomial exp2, 16-bit packing, phase tracing.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#ifdef RD_XATTN_TRACE
// phase cycle counters of the first warp of each softmax warpgroup (summed over CTAs), read by tools/probes/xattn_bench.cu;
// every kernel file defines its own `rd_xattn_trace[16]` in its namespace
#define RD_TRACE_DECL long long tr_t = clock64(); unsigned long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define RD_TRACE(i) { const long long tr_n = clock64(); tr_acc[i] += (unsigned long long)(tr_n - tr_t); tr_t = tr_n; }
#define RD_TRACE_FLUSH if ((threadIdx.x & 127) == 0) { const int tr_o = (threadIdx.x >> 7) * 8; for (int tr_i = 0; tr_i < 7; ++tr_i) atomicAdd(&rd_xattn_trace[tr_i], tr_acc[tr_i]); atomicAdd(&rd_xattn_trace[8 + (tr_o >> 3)], 1ull); }
#else
#define RD_TRACE_DECL
#define RD_TRACE(i)
#define RD_TRACE_FLUSH
#endif

namespace rd {
namespace xa {

template <typename T> struct PackPair;
template <> struct PackPair<__half> {
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
};
template <> struct PackPair<__nv_bfloat16> {
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));   // FMNMX3 on sm_100
    return y;
}
// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2, two results per issue slot)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
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
// 2^x for two values on the FMA/ALU pipes (the MUFU pipe, 16 results/clk/SM, is the first ceiling of head_dim-64
// attention): round-to-nearest split x = xi + xf, xf in [-0.5, 0.5], degree-3 minimax polynomial for 2^xf (max
// relative error 7.5e-5, below the fp16/bf16 rounding of P), exponent patched in with one integer multiply-add.
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& p0, float& p1) {
    float x0, x1;
    unpack2(x2, x0, x1);
    x0 = fmaxf(x0, -125.0f);
    x1 = fmaxf(x1, -125.0f);
    const uint64_t x = pack2(x0, x1);
    const uint64_t magic = pack2(12582912.0f, 12582912.0f);     // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const uint64_t t = fadd2(x, magic);
    const uint64_t xf = fsub2(x, fsub2(t, magic));
    uint64_t p = ffma2(pack2(0.05517164617776871f, 0.05517164617776871f), xf, pack2(0.2426111251115799f, 0.2426111251115799f));
    p = ffma2(p, xf, pack2(0.6932609677314758f, 0.6932609677314758f));
    p = ffma2(p, xf, pack2(0.9999280571937561f, 0.9999280571937561f));
    float t0, t1, q0, q1;
    unpack2(t, t0, t1);
    unpack2(p, q0, q1);
    p0 = __int_as_float(__float_as_int(t0) * (1 << 23) + __float_as_int(q0));
    p1 = __int_as_float(__float_as_int(t1) * (1 << 23) + __float_as_int(q1));
}

}  // namespace xa
}  // namespace rd
