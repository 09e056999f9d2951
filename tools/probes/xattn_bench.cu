// Stand-alone harness for the cross-frame attention kernel: no Python, no torch (fast turn-around on a GPU box).
// Times rd_xattn_fwd on the strided [B, S, H*64] layout, checks sampled rows against a double-precision CPU
// softmax(QK^T)V, and (with -DRD_XATTN_TRACE) prints the per-phase cycle budget of the softmax warps.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo [-DRD_XATTN_TRACE] \
//        -I include -I rollingdepth_b200/csrc tools/probes/xattn_bench.cu rollingdepth_b200/csrc/rd_abi.cu \
//        rollingdepth_b200/csrc/xattn_f32.cu rollingdepth_b200/csrc/xattn_sm100_q3.cu \
//        -o tools/probes/bin/xattn_bench
//   xattn_bench [B H S [flags [iters [bf16]]]]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

// archived variants (build with -I rollingdepth_b200/csrc; see attn_variants/README.md for their measurements)
#if defined(XATTN_V6)
#include "attn_variants/xattn_v6_64key_steps_double_buffered_prefetch.cu.txt"
#elif defined(XATTN_V7)
#include "attn_variants/xattn_v7_four_tiles_p_aliased_polling_issuer.cu.txt"
#else
#include "../../rollingdepth_b200/csrc/xattn_sm100.cu"
#endif

template <typename T>
__global__ void fill_kernel(T* p, size_t n, uint32_t seed, float scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i * 2654435761u ^ seed;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        uint32_t y = x * 0x9e3779b9u + 0x85ebca6bu;
        y ^= y >> 15; y *= 0x2c1b3c6du; y ^= y >> 12;
        // sum of uniforms ~ roughly normal, unit variance
        const float u = ((x & 0xffff) + (x >> 16) + (y & 0xffff) + (y >> 16)) * (1.0f / 65536.0f) - 2.0f;
        p[i] = rd::Cvt<T>::from_f(u * 1.7320508f * scale);
    }
}

template <typename T>
static float to_float(T v);
template <> float to_float<__half>(__half v) { return __half2float(v); }
template <> float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
int run(int B, int H, int S, int flags, int iters, rd_dtype dt) {
    const size_t n = (size_t)B * S * H * 64;
    T *q, *k, *v, *o;
    cudaMalloc(&q, n * 2); cudaMalloc(&k, n * 2); cudaMalloc(&v, n * 2); cudaMalloc(&o, n * 2);
    fill_kernel<<<1184, 256>>>(q, n, 1u, 1.0f);
    fill_kernel<<<1184, 256>>>(k, n, 2u, 1.0f);
    fill_kernel<<<1184, 256>>>(v, n, 3u, 1.0f);
    cudaMemset(o, 0, n * 2);
    const int64_t sb = (int64_t)S * H * 64, ss = (int64_t)H * 64, sh = 64;
    const float scale = 0.125f;
    auto launch = [&]() {
        return rd_xattn_fwd(q, k, v, o, B, H, S, S, 64, sb, ss, sh, sb, ss, sh, sb, ss, sh, sb, ss, sh, dt, scale, flags, nullptr);
    };
    for (int i = 0; i < 3; ++i)
        if (int e = launch()) { printf("launch failed %d: %s\n", e, rd_last_error()); return 1; }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
#ifdef RD_XATTN_TRACE
    unsigned long long zeros[16] = {0};
    cudaMemcpyToSymbol(rd::rd_xattn_trace, zeros, sizeof(zeros));
#endif
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    const double flops = 4.0 * B * H * (double)S * S * 64;
    printf("B=%d H=%d S=%d %s flags=%d: %.4f ms  %.1f TFLOP/s\n", B, H, S, dt == RD_F16 ? "f16" : "bf16", flags, ms, flops / ms / 1e9);
#if defined(RD_XATTN_TRACE)
    unsigned long long tr[16];
    cudaMemcpyFromSymbol(tr, rd::rd_xattn_trace, sizeof(tr));
#if defined(XATTN_V6) || defined(XATTN_V7)
    const int step_keys = 64;
    const char* names[8] = {"wait S_FULL", "LDTM + wait", "max + rescale", "exp + pack + STTM", "st wait + P_FULL", "-", "-", "-"};
#else
    const int step_keys = 128;
    const char* names[8] = {"wait S_FULL", "LDTM S + S_FREE", "mask + row max", "wait O_READY", "rescale (lazy)", "exp + pack + STTM", "st wait + P_FULL", "-"};
#endif
    const double tiles = (double)(tr[8] + tr[9] + tr[10]) * ((S + step_keys - 1) / step_keys);   // steps x traced warps
    double tot = 0;
    for (int i = 0; i < 7; ++i) tot += tr[i] / tiles;
    for (int i = 0; i < 7; ++i) printf("  %-26s %8.1f cycles/step  (%4.1f%%)\n", names[i], tr[i] / tiles, 100.0 * tr[i] / tiles / tot);
    printf("  total                      %8.1f cycles per %d-key step per softmax warp\n", tot, step_keys);
#ifdef RD_XATTN_TIMELINE
    {
        static long long tl[4][32][8];
        cudaMemcpyFromSymbol(tl, rd::rd_xattn_tl, sizeof(tl));
        const long long t0 = tl[0][0][0];
        printf("  timeline of CTA 0 (cycles relative to softmax warp 0, step 64, point 0)\n");
        printf("  step |  sm0: ldtm-issued  p-free  exp-done  max-done  p_full  resc |  sm4: same | iss0: start v_full p_full pv-issued qk-issued | iss1: same\n");
        for (int j = 0; j < 24; ++j) {
            printf("  %4d |", 64 + j);
            for (int a = 0; a < 2; ++a) { for (int i = 0; i < 6; ++i) printf(" %6lld", tl[a][j][i] - t0); printf(" |"); }
            for (int a = 2; a < 4; ++a) { for (int i = 0; i < 5; ++i) printf(" %6lld", tl[a][j][i] - t0); printf(" |"); }
            printf("\n");
        }
    }
#endif
#endif
    // sampled check
    std::vector<T> hq(n), hk(n), hv(n), ho(n);
    cudaMemcpy(hq.data(), q, n * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(hk.data(), k, n * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(hv.data(), v, n * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho.data(), o, n * 2, cudaMemcpyDeviceToHost);
    double max_err = 0;
    const int n_check = 24;
    std::vector<double> sc(S);
    for (int c = 0; c < n_check; ++c) {
        const int b = c % B, h = (c * 7) % H;
        int r = (int)(((long long)c * 2654435761ll) % S);
        if (c == 0) r = 0;
        if (c == 1) r = S - 1;
        const T* qr = &hq[(size_t)b * sb + (size_t)r * ss + h * sh];
        double mx = -1e300;
        for (int j = 0; j < S; ++j) {
            const T* kr = &hk[(size_t)b * sb + (size_t)j * ss + h * sh];
            double d = 0;
            for (int e = 0; e < 64; ++e) d += (double)to_float(qr[e]) * to_float(kr[e]);
            sc[j] = d * scale;
            mx = std::max(mx, sc[j]);
        }
        double l = 0, acc[64] = {0};
        for (int j = 0; j < S; ++j) {
            const double p = std::exp(sc[j] - mx);
            l += p;
            const T* vr = &hv[(size_t)b * sb + (size_t)j * ss + h * sh];
            for (int e = 0; e < 64; ++e) acc[e] += p * to_float(vr[e]);
        }
        const T* orow = &ho[(size_t)b * sb + (size_t)r * ss + h * sh];
        for (int e = 0; e < 64; ++e) max_err = std::max(max_err, std::fabs(acc[e] / l - (double)to_float(orow[e])));
    }
    printf("  max-abs error on %d sampled rows: %.3e %s\n", n_check, max_err, max_err < 2e-2 ? "OK" : "FAIL");
    cudaFree(q); cudaFree(k); cudaFree(v); cudaFree(o);
    return max_err < 2e-2 ? 0 : 2;
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 1;
    const int H = argc > 2 ? atoi(argv[2]) : 5;
    const int S = argc > 3 ? atoi(argv[3]) : 27648;
    const int flags = argc > 4 ? atoi(argv[4]) : 0;
    const int iters = argc > 5 ? atoi(argv[5]) : 10;
    const bool bf16 = argc > 6 && atoi(argv[6]) != 0;
    return bf16 ? run<__nv_bfloat16>(B, H, S, flags, iters, RD_BF16) : run<__half>(B, H, S, flags, iters, RD_F16);
}
