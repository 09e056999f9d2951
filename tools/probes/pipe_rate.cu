// Probe: issue cost (cycles per warp-instruction per SM sub-partition) of the instructions the softmax inner loop is
// made of, alone and mixed with MUFU.EX2, for 1 and 2 resident warps per sub-partition.  16 independent dependency
// chains per thread so that latency is covered.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define N 16

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

enum { FFMA_REG, FFMA_IMM, FADD_REG, FADD_IMM, FFMA2_, FADD2_, FMUL2_, F2FP_, FMNMX3_, LEA_, IMAD_IMM, MUFU_, MUFU_FFMA2_FADD2_F2FP, MUFU_FFMA_FADD_F2FP,
       POLY_SCALAR, POLY_PACKED, HFMA2_, N_MODES };
static const char* names[N_MODES] = {"FFMA 3-reg", "FFMA imm addend", "FADD reg", "FADD imm", "FFMA2 (f32x2)", "FADD2 (f32x2)", "FMUL2 (f32x2)", "F2FP pack",
    "FMNMX3", "LEA (shl+add)", "IMAD imm", "MUFU.EX2", "mix/pair: 2 MUFU + FFMA2 + FADD2 + F2FP", "mix/pair: 2 MUFU + 2 FFMA + 2 FADD + F2FP",
    "poly exp2 scalar (per element)", "poly exp2 packed (per pair)", "HFMA2"};

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc, float c0, float c1) {
    float a[N];
    uint64_t p[N];
    uint32_t u[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        a[i] = -1.0f - 0.01f * (threadIdx.x + i);
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] * 0.5f));
        u[i] = threadIdx.x * 77 + i;
    }
    uint64_t c0p, c1p;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c0p) : "f"(c0));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(c1p) : "f"(c1));
    float s = 0.f;
    uint64_t s2 = 0;
    unsigned pk = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE == FFMA_REG) a[i] = fmaf(a[i], c0, c1);
            if (MODE == FFMA_IMM) a[i] = fmaf(a[i], c0, 0.2426111251115799f);
            if (MODE == FADD_REG) a[i] = a[i] + c1;
            if (MODE == FADD_IMM) a[i] = a[i] + 12582912.0f;
            if (MODE == FFMA2_) p[i] = ffma2(p[i], c0p, c1p);
            if (MODE == FADD2_) p[i] = fadd2(p[i], c1p);
            if (MODE == FMUL2_) p[i] = fmul2(p[i], c0p);
            if (MODE == F2FP_) { __half2 h = __floats2half2_rn(a[i], c1); u[i] = *reinterpret_cast<unsigned*>(&h); a[i] = __uint_as_float(u[i]); }
            if (MODE == FMNMX3_) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1));
            if (MODE == LEA_) asm volatile("{.reg .b32 t; shl.b32 t, %0, 23; add.s32 %0, t, %1;}" : "+r"(u[i]) : "r"(__float_as_uint(c0)));
            if (MODE == IMAD_IMM) asm volatile("mad.lo.u32 %0, %0, 8388608, %1;" : "+r"(u[i]) : "r"(__float_as_uint(c0)));
            if (MODE == MUFU_) a[i] = ex2(a[i]);
            if (MODE == HFMA2_) { __half2 h = *reinterpret_cast<__half2*>(&u[i]); __half2 cc = *reinterpret_cast<const __half2*>(&c0); h = __hfma2(h, cc, cc); u[i] = *reinterpret_cast<unsigned*>(&h); }
        }
        if (MODE == MUFU_FFMA2_FADD2_F2FP) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const uint64_t x2 = ffma2(p[i], c0p, c1p);
                float x0, x1;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x2));
                const float p0 = ex2(x0), p1 = ex2(x1);
                uint64_t pp;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pp) : "f"(p0), "f"(p1));
                s2 = fadd2(s2, pp);
                __half2 h = __floats2half2_rn(p0, p1);
                pk ^= *reinterpret_cast<unsigned*>(&h);
                p[i] = x2;
            }
        }
        if (MODE == MUFU_FFMA_FADD_F2FP) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float x0 = fmaf(a[i], c0, c1), x1 = fmaf(a[i], c1, c0);
                const float p0 = ex2(x0), p1 = ex2(x1);
                s += p0; s += p1;
                __half2 h = __floats2half2_rn(p0, p1);
                pk ^= *reinterpret_cast<unsigned*>(&h);
                a[i] = x0;
            }
        }
        if (MODE == POLY_SCALAR) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float x = fmaxf(fmaf(a[i], c0, c1), -125.0f);
                const float t = x + 12582912.0f;
                const float xf = x - (t - 12582912.0f);
                float q = fmaf(xf, 0.05517164617776871f, 0.2426111251115799f);
                q = fmaf(q, xf, 0.6932609677314758f);
                q = fmaf(q, xf, 0.9999280571937561f);
                const float r = __uint_as_float((__float_as_uint(t) << 23) + __float_as_uint(q));
                s += r;
                a[i] = x;
            }
        }
        if (MODE == POLY_PACKED) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const uint64_t x = ffma2(p[i], c0p, c1p);
                uint64_t magic, k3, k2, k1, k0;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(magic) : "f"(12582912.0f));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(k3) : "f"(0.05517164617776871f));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(k2) : "f"(0.2426111251115799f));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(k1) : "f"(0.6932609677314758f));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(k0) : "f"(0.9999280571937561f));
                const uint64_t t = fadd2(x, magic);
                uint64_t nm;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(nm) : "f"(-12582912.0f));
                const uint64_t xi = fadd2(t, nm);
                uint64_t nxi = xi ^ 0x8000000080000000ull;
                const uint64_t xf = fadd2(x, nxi);
                uint64_t q = ffma2(k3, xf, k2);
                q = ffma2(q, xf, k1);
                q = ffma2(q, xf, k0);
                float t0f, t1f, q0, q1;
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(t0f), "=f"(t1f) : "l"(t));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
                const float r0 = __uint_as_float((__float_as_uint(t0f) << 23) + __float_as_uint(q0));
                const float r1 = __uint_as_float((__float_as_uint(t1f) << 23) + __float_as_uint(q1));
                s += r0; s += r1;
                p[i] = x;
            }
        }
    }
    long long t1 = clock64();
    float r = s + __uint_as_float(pk);
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s2));
    r += lo + hi;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        r += a[i] + lo + hi + __uint_as_float(u[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(float* out, long long* cyc) {
    const int iters = 4000;
    for (int warps : {4, 8, 16}) {
        long long h = 0;
        k<MODE><<<148, warps * 32>>>(out, iters, cyc, 0.999f, -0.001f);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        // cycles per (loop element) per sub-partition: total elements per SMSP = iters * N * warps / 4
        const double per = (double)h / ((double)iters * N * warps / 4.0);
        printf("%-46s warps/SMSP %d: %6.2f cycles per element per sub-partition\n", names[MODE], warps / 4, per);
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 2048 * 4); cudaMalloc(&cyc, 8);
    run<FFMA_REG>(out, cyc); run<FFMA_IMM>(out, cyc); run<FADD_REG>(out, cyc); run<FADD_IMM>(out, cyc);
    run<FFMA2_>(out, cyc); run<FADD2_>(out, cyc); run<FMUL2_>(out, cyc); run<F2FP_>(out, cyc); run<FMNMX3_>(out, cyc);
    run<LEA_>(out, cyc); run<IMAD_IMM>(out, cyc); run<HFMA2_>(out, cyc); run<MUFU_>(out, cyc);
    run<MUFU_FFMA2_FADD2_F2FP>(out, cyc); run<MUFU_FFMA_FADD_F2FP>(out, cyc);
    run<POLY_SCALAR>(out, cyc); run<POLY_PACKED>(out, cyc);
    return 0;
}
