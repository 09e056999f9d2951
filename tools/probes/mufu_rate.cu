// Probe: sustained MUFU.EX2 throughput per SM (results/clk) as a function of resident warps, plus a mixed
// FFMA+EX2+FADD+F2FP stream shaped like the softmax inner loop.  Build: nvcc -arch=sm_100a -O3 mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = -1.0f - 0.01f * (threadIdx.x + i);
    float s = 0.f; unsigned pk = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (MODE == 0) { a[i] = ex2(a[i]); a[i + 1] = ex2(a[i + 1]); }
            else {
                const float x0 = fmaf(a[i], 0.18f, -0.5f), x1 = fmaf(a[i + 1], 0.18f, -0.5f);
                const float p0 = ex2(x0), p1 = ex2(x1);
                s += p0 + p1;
                __half2 h = __floats2half2_rn(p0, p1);
                pk ^= *reinterpret_cast<unsigned*>(&h);
                a[i] -= 1e-3f; a[i + 1] -= 1e-3f;
            }
        }
    }
    long long t1 = clock64();
    float r = s + __uint_as_float(pk);
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 2048 * 4); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            long long h = 0;
            if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double ex = (double)iters * 16 * warps * 32;
            printf("mode %d warps/SM %2d: %.2f ex2 per clk per SM (cycles %lld)\n", mode, warps, ex / (double)h, h);
        }
    return 0;
}
