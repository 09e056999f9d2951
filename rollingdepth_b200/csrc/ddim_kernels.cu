// DDIM step / add_noise as single element-wise launches (sm_100a).
// Replaces diffusers/src/diffusers/schedulers/scheduling_ddim.py:342-468 (eta = 0) and :471-495.
// Every reference op rounds its result to the tensor dtype; the same roundings are applied here
// so that fp16 runs reproduce the reference's op-by-op arithmetic bit for bit (__fmul_rn/__fadd_rn
// are never contracted into FMAs, so the fp32 path keeps the per-op roundings too).
#include "common.cuh"

namespace rd {

template <typename T>
__global__ void __launch_bounds__(256) ddim_step_kernel(const T* __restrict__ mo, const T* __restrict__ sm, T* __restrict__ out,
                                                        long long n, float sa, float sb, float sap, float sbp, int ptype,
                                                        float clip) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = Cvt<T>::to_f(mo[i]);
        const float x = Cvt<T>::to_f(sm[i]);
        float x0, eps;
        if (ptype == 2) {          // v_prediction (:421-423)
            x0 = Cvt<T>::round(__fsub_rn(Cvt<T>::round(__fmul_rn(sa, x)), Cvt<T>::round(__fmul_rn(sb, v))));
            eps = Cvt<T>::round(__fadd_rn(Cvt<T>::round(__fmul_rn(sa, v)), Cvt<T>::round(__fmul_rn(sb, x))));
        } else if (ptype == 0) {   // epsilon (:415-417)
            // torch's CUDA true-divide by a host scalar multiplies by the fp32 reciprocal
            x0 = Cvt<T>::round(__fmul_rn(Cvt<T>::round(__fsub_rn(x, Cvt<T>::round(__fmul_rn(sb, v)))), __frcp_rn(sa)));
            eps = v;
        } else {                   // sample (:418-420)
            x0 = v;
            eps = Cvt<T>::round(__fmul_rn(Cvt<T>::round(__fsub_rn(x, Cvt<T>::round(__fmul_rn(sa, x0)))), __frcp_rn(sb)));
        }
        if (clip > 0.f) x0 = fminf(fmaxf(x0, -clip), clip);   // :433-436
        const float dir = Cvt<T>::round(__fmul_rn(sbp, eps));                            // :449
        out[i] = Cvt<T>::from_f(__fadd_rn(Cvt<T>::round(__fmul_rn(sap, x0)), dir));      // :452
    }
}

template <typename T>
__global__ void __launch_bounds__(256) ddim_add_noise_kernel(const T* __restrict__ x, const T* __restrict__ nz, T* __restrict__ out,
                                                             long long n, float sa, float sb) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = Cvt<T>::round(__fmul_rn(sa, Cvt<T>::to_f(x[i])));
        const float b = Cvt<T>::round(__fmul_rn(sb, Cvt<T>::to_f(nz[i])));
        out[i] = Cvt<T>::from_f(__fadd_rn(a, b));
    }
}

static int grid_for(long long n) {
    long long want = (n + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    return (int)(want > cap ? cap : (want < 1 ? 1 : want));
}

}  // namespace rd

using namespace rd;

extern "C" {

int rd_ddim_step(const void* model_out, const void* sample, void* prev_sample, int64_t n, float sqrt_alpha_t,
                 float sqrt_beta_t, float sqrt_alpha_prev, float sqrt_one_minus_alpha_prev, int prediction_type,
                 float clip, rd_dtype dtype, void* stream) {
    RD_CHECK_ARG(model_out && sample && prev_sample && n > 0, "bad pointer / size");
    RD_CHECK_ARG(prediction_type >= 0 && prediction_type <= 2, "prediction_type %d not in {0,1,2}", prediction_type);
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n);
    switch (dtype) {
        case RD_F32:
            ddim_step_kernel<float><<<g, 256, 0, st>>>((const float*)model_out, (const float*)sample, (float*)prev_sample, n,
                                                      sqrt_alpha_t, sqrt_beta_t, sqrt_alpha_prev, sqrt_one_minus_alpha_prev,
                                                      prediction_type, clip);
            break;
        case RD_F16:
            ddim_step_kernel<__half><<<g, 256, 0, st>>>((const __half*)model_out, (const __half*)sample, (__half*)prev_sample, n,
                                                       sqrt_alpha_t, sqrt_beta_t, sqrt_alpha_prev, sqrt_one_minus_alpha_prev,
                                                       prediction_type, clip);
            break;
        case RD_BF16:
            ddim_step_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)model_out, (const __nv_bfloat16*)sample,
                                                              (__nv_bfloat16*)prev_sample, n, sqrt_alpha_t, sqrt_beta_t,
                                                              sqrt_alpha_prev, sqrt_one_minus_alpha_prev, prediction_type, clip);
            break;
        default: set_error("bad dtype"); return RD_ERR_BAD_ARG;
    }
    RD_LAUNCH_CHECK("ddim_step_kernel");
    return RD_OK;
}

int rd_ddim_add_noise(const void* original, const void* noise, void* out, int64_t n, float sqrt_alpha_prod,
                      float sqrt_one_minus_alpha_prod, rd_dtype dtype, void* stream) {
    RD_CHECK_ARG(original && noise && out && n > 0, "bad pointer / size");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n);
    switch (dtype) {
        case RD_F32:
            ddim_add_noise_kernel<float><<<g, 256, 0, st>>>((const float*)original, (const float*)noise, (float*)out, n,
                                                           sqrt_alpha_prod, sqrt_one_minus_alpha_prod);
            break;
        case RD_F16:
            ddim_add_noise_kernel<__half><<<g, 256, 0, st>>>((const __half*)original, (const __half*)noise, (__half*)out, n,
                                                            sqrt_alpha_prod, sqrt_one_minus_alpha_prod);
            break;
        case RD_BF16:
            ddim_add_noise_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)original, (const __nv_bfloat16*)noise,
                                                                   (__nv_bfloat16*)out, n, sqrt_alpha_prod,
                                                                   sqrt_one_minus_alpha_prod);
            break;
        default: set_error("bad dtype"); return RD_ERR_BAD_ARG;
    }
    RD_LAUNCH_CHECK("ddim_add_noise_kernel");
    return RD_OK;
}

}  // extern "C"
