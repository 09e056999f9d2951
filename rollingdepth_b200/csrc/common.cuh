// Shared helpers for the librd_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rd_b200.h"

namespace rd {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define RD_CHECK_ARG(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            rd::set_error(__VA_ARGS__);                 \
            return RD_ERR_BAD_ARG;                      \
        }                                               \
    } while (0)

#define RD_CUDA(call)                                               \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return rd::cuda_fail(_e, #call);     \
    } while (0)

#define RD_LAUNCH_CHECK(what)                                       \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return rd::cuda_fail(_e, what);      \
    } while (0)

int num_sms();            // SM count of the CURRENT device (cached per device ordinal)
int current_device();     // cudaGetDevice; -1 on error

// One-time per-DEVICE configuration of a kernel (cudaFuncSetAttribute is per device, not per process):
//   static PerDeviceOnce once;  if (once.needed()) { RD_CUDA(cudaFuncSetAttribute(...)); once.done(); }
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
    bool flag[kMaxDevices] = {};
    bool needed() const { const int d = current_device(); return d < 0 || d >= kMaxDevices || !flag[d]; }
    void done() { const int d = current_device(); if (d >= 0 && d < kMaxDevices) flag[d] = true; }
};

// ---------------------------------------------------------------- dtype helpers
template <typename T> struct Cvt;
template <> struct Cvt<float> {
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
    static __device__ __forceinline__ float round(float v) { return v; }  // round-trip through T
};
template <> struct Cvt<__half> {
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
    static __device__ __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
};
template <> struct Cvt<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
    static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 16-byte load that does not allocate in L1 (data is touched once per launch)
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream4(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

}  // namespace rd
