// Cross-frame self-attention forward for sm_100a: flash-style, head_dim 64, fp16/bf16 in,
// fp32 accumulate.  Replaces the fused-attention library call of the reference's processors
// (diffusers/src/diffusers/models/attention_processor.py:2251-2253 and :2036-2038).
//
// One CTA = one 128-row query tile of one (batch, head).  Warp roles:
//   warps 0-3  softmax: each thread owns one query row (TMEM lane); S -> registers via
//              tcgen05.ld, online softmax with lazy (thresholded) rescaling, P written back to
//              TMEM (fp16/bf16 packed) with tcgen05.st, O rescaled in TMEM when needed, epilogue.
//   warp 4     TMA producer: Q once, then K/V tiles of 128 keys through two 2-stage rings.
//   warp 5     tcgen05.mma issuer (one elected thread): S = Q K^T (SS mode, both K-major,
//              SWIZZLE_128B), O += P V (A = P from TMEM, B = V MN-major from shared memory).
// TMEM columns: S [0,128) fp32, O [128,192) fp32, P [192,256) packed 16-bit  => 256 columns,
// two CTAs per SM.  QK^T of tile j+1 is issued as soon as the softmax warps have pulled S_j into
// registers, so the tensor pipe works on the next scores while the exponentials are computed.
#include <cuda.h>
#include <math.h>
#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.cuh"

namespace rd {

using namespace ptx;

constexpr int kBM = 128;       // query rows per CTA
constexpr int kBN = 128;       // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kStages = 2;
constexpr int kTileBytes = kBN * kD * 2;   // 16384
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColO = 128, kColP = 192;
constexpr float kRescaleThreshold = 8.0f;  // log2 units; p stays <= 2^8

struct XattnSmem {
    // offsets inside the 1024-byte aligned dynamic shared memory block
    static constexpr int q = 0;
    static constexpr int k = q + kTileBytes;
    static constexpr int v = k + kStages * kTileBytes;
    static constexpr int bars = v + kStages * kTileBytes;       // 16 x 8 bytes
    static constexpr int tmem_ptr = bars + 16 * 8;
    static constexpr int p = ((tmem_ptr + 16 + 1023) / 1024) * 1024;   // only in the P-via-smem variant
    static constexpr int bytes_tmem_variant = tmem_ptr + 16;
    static constexpr int bytes_smem_variant = p + 2 * kTileBytes;
};

enum BarId { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_FREE = 10, P_FULL = 11, O_READY = 12 };

struct XattnParams {
    void* o;
    long long o_sb, o_ss, o_sh;
    int Sq, Sk, n_kv_tiles;
    float scale_log2;   // scale * log2(e)
};

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
// 2^x on the FMA/ALU pipes (the MUFU pipe, 16 results/clk/SM, is the first ceiling of head_dim-64 attention):
// round-to-nearest split x = xi + xf, xf in [-0.5, 0.5], degree-3 minimax polynomial for 2^xf
// (max relative error 7.5e-5, below the fp16/bf16 rounding of P), exponent patched in with integer arithmetic.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;            // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float xf = x - (t - 12582912.0f);
    float p = fmaf(0.05517164617776871f, xf, 0.2426111251115799f);
    p = fmaf(p, xf, 0.6932609677314758f);
    p = fmaf(p, xf, 0.9999280571937561f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// POLY: every POLY-th score pair takes the polynomial exp2 path (0 = all on MUFU)
template <typename T, bool P_SMEM, int POLY>
__global__ void __launch_bounds__(kThreads, P_SMEM ? 1 : 2)
xattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const XattnParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XattnSmem::bars);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + XattnSmem::tmem_ptr);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_tiles = prm.n_kv_tiles;

    if (threadIdx.x == 0) {
        mbar_init(&bars[Q_FULL], 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars[K_FULL + s], 1);
            mbar_init(&bars[K_EMPTY + s], 1);
            mbar_init(&bars[V_FULL + s], 1);
            mbar_init(&bars[V_EMPTY + s], 1);
        }
        mbar_init(&bars[S_FULL], 1);
        mbar_init(&bars[S_FREE], 128);
        mbar_init(&bars[P_FULL], 128);
        mbar_init(&bars[O_READY], 1);
        mbar_fence_init();
    }
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
    }
    if (warp == 5) {
        tmem_alloc(tmem_ptr_smem, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 4) {
        // ===================================================== TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars[Q_FULL], kTileBytes);
            tma_load_4d(smem + XattnSmem::q, &tm_q, &bars[Q_FULL], 0, q_tile * kBM, head, batch);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % kStages;
                const uint32_t ph = (j / kStages) & 1;
                mbar_wait(&bars[K_EMPTY + st], ph ^ 1);
                mbar_arrive_expect_tx(&bars[K_FULL + st], kTileBytes);
                tma_load_4d(smem + XattnSmem::k + st * kTileBytes, &tm_k, &bars[K_FULL + st], 0, j * kBN, head, batch);
                mbar_wait(&bars[V_EMPTY + st], ph ^ 1);
                mbar_arrive_expect_tx(&bars[V_FULL + st], kTileBytes);
                tma_load_4d(smem + XattnSmem::v + st * kTileBytes, &tm_v, &bars[V_FULL + st], 0, j * kBN, head, batch);
            }
        }
    } else if (warp == 5) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t fmt = sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            constexpr uint32_t idesc_qk = make_idesc_f16(fmt, 0, 0, kBM, kBN);
            constexpr uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, kBM, kD);   // B = V is MN-major
            const uint32_t tS = tmem_base + kColS, tO = tmem_base + kColO, tP = tmem_base + kColP;
            const uint64_t dq = make_smem_desc_sw128(smem_u32(smem + XattnSmem::q), 16, 1024);

            auto issue_qk = [&](int j) {
                const int st = j % kStages;
                mbar_wait(&bars[K_FULL + st], (j / kStages) & 1);
                if (j > 0) mbar_wait(&bars[S_FREE], (j - 1) & 1);   // softmax(j-1) has S in registers
                tc_fence_after();
                const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + XattnSmem::k + st * kTileBytes), 16, 1024);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)                 // +32 bytes along K inside the 128B swizzle row
                    mma_f16_ss(tS, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0);
                mma_commit(&bars[K_EMPTY + st]);
                mma_commit(&bars[S_FULL]);
            };

            mbar_wait(&bars[Q_FULL], 0);
            issue_qk(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_qk(j + 1);
                const int st = j % kStages;
                mbar_wait(&bars[V_FULL + st], (j / kStages) & 1);
                mbar_wait(&bars[P_FULL], j & 1);
                tc_fence_after();
                const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + XattnSmem::v + st * kTileBytes), 16, 1024);
#pragma unroll
                for (int kk = 0; kk < kBN / 16; ++kk) {
                    const uint32_t acc = (j > 0 || kk > 0) ? 1u : 0u;
                    // V advances 16 key rows = 2048 bytes per K step
                    if constexpr (P_SMEM) {
                        const uint64_t dp = make_smem_desc_sw128(
                            smem_u32(smem + XattnSmem::p + (kk / 4) * kTileBytes) + (kk % 4) * 32, 16, 1024);
                        mma_f16_ss(tO, dp, dv + (uint64_t)(128 * kk), idesc_pv, acc);
                    } else {
                        mma_f16_ts(tO, tP + 8 * kk, dv + (uint64_t)(128 * kk), idesc_pv, acc);
                    }
                }
                mma_commit(&bars[V_EMPTY + st]);
                mma_commit(&bars[O_READY]);
            }
        }
    } else {
        // ===================================================== softmax warps (one query row per thread)
        const int row = warp * 32 + lane;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + kColS, tO = tmem_base + lane_off + kColO, tP = tmem_base + lane_off + kColP;
        const float sl2 = prm.scale_log2;
        float m_used = -INFINITY;   // raw-score units
        float l = 0.f;

        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&bars[S_FULL], j & 1);
            tc_fence_after();
            uint32_t sr[4][32];
            tmem_ld32(tS + 0, sr[0]);
            tmem_ld32(tS + 32, sr[1]);
            tmem_ld32(tS + 64, sr[2]);
            tmem_ld32(tS + 96, sr[3]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars[S_FREE]);

            const int valid = prm.Sk - j * kBN;   // keys of this tile that exist
            if (valid < kBN) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i >= valid) sr[c][i] = 0xff800000u;   // -inf
            }
            // row maximum: 8 independent chains of 3-input max (a single dependent chain of 128 FMNMX is ~500 cycles)
            float mx[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) mx[q] = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                    mx[(i >> 1) & 7] = max3(mx[(i >> 1) & 7], __uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1]));
            const float mt = fmaxf(max3(mx[0], mx[1], mx[2]), max3(max3(mx[3], mx[4], mx[5]), mx[6], mx[7]));
            const float m_new = fmaxf(m_used, mt);

            if (j > 0) {
                // P (and O) may only be touched once PV(j-1) has completed
                mbar_wait(&bars[O_READY], (j - 1) & 1);
                tc_fence_after();
                const bool need = (m_new - m_used) * sl2 > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    const float f = ex2((m_used - m_new) * sl2);
                    l *= f;
                    uint32_t orr[2][32];
                    tmem_ld32(tO + 0, orr[0]);
                    tmem_ld32(tO + 32, orr[1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i) orr[c][i] = __float_as_uint(__uint_as_float(orr[c][i]) * f);
                    tmem_st32(tO + 0, orr[0]);
                    tmem_st32(tO + 32, orr[1]);
                    tmem_st_wait();
                    m_used = m_new;
                }
            } else {
                m_used = m_new;
            }

            const float mb = m_used * sl2;
            uint32_t pk[2][32];
            float ls[4] = {0.f, 0.f, 0.f, 0.f};       // independent partial row sums
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float x0 = fmaf(__uint_as_float(sr[c][i]), sl2, -mb);
                    const float x1 = fmaf(__uint_as_float(sr[c][i + 1]), sl2, -mb);
                    float p0, p1;
                    if (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == POLY - 1) {
                        p0 = ex2_poly(x0);
                        p1 = ex2_poly(x1);
                    } else {
                        p0 = ex2(x0);
                        p1 = ex2(x1);
                    }
                    ls[(i >> 1) & 3] += p0 + p1;
                    pk[c >> 1][(c & 1) * 16 + (i >> 1)] = PackPair<T>::pack(p0, p1);
                }
            }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
            if constexpr (P_SMEM) {
                // K-major SWIZZLE_128B: two [128 x 64] sub-tiles, 16-byte chunk index XOR (row % 8)
                uint8_t* prow = smem + XattnSmem::p + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
                for (int c16 = 0; c16 < 16; ++c16) {
                    const int sub = c16 >> 3, cc = c16 & 7;
                    uint4 val;
                    const uint32_t* src = &pk[c16 >> 3][(c16 & 7) * 4];
                    val.x = src[0]; val.y = src[1]; val.z = src[2]; val.w = src[3];
                    *reinterpret_cast<uint4*>(prow + sub * kTileBytes + ((cc ^ (row & 7)) << 4)) = val;
                }
                fence_proxy_async_smem();
            } else {
                tmem_st32(tP + 0, pk[0]);
                tmem_st32(tP + 32, pk[1]);
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(&bars[P_FULL]);
        }

        // ------------------------------------------------- epilogue: O / l -> global
        mbar_wait(&bars[O_READY], (n_tiles - 1) & 1);
        tc_fence_after();
        uint32_t orr[2][32];
        tmem_ld32(tO + 0, orr[0]);
        tmem_ld32(tO + 32, orr[1]);
        tmem_ld_wait();
        const float inv_l = 1.0f / l;
        const int grow = q_tile * kBM + row;
        if (grow < prm.Sq) {
            T* dst = reinterpret_cast<T*>(prm.o) + (long long)batch * prm.o_sb + (long long)grow * prm.o_ss +
                     (long long)head * prm.o_sh;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 val;
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = c * 8 + i * 2;
                    w[i] = PackPair<T>::pack(__uint_as_float(orr[e >> 5][e & 31]) * inv_l,
                                             __uint_as_float(orr[(e + 1) >> 5][(e + 1) & 31]) * inv_l);
                }
                val.x = w[0]; val.y = w[1]; val.z = w[2]; val.w = w[3];
                *reinterpret_cast<uint4*>(dst + c * 8) = val;
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side (tensor maps: tensormap.cuh)
// ---------------------------------------------------------------------------------------------
int xattn_f32_launch(const float* q, const float* k, const float* v, float* o, int B, int H, int Sq, int Sk,
                     const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, float scale,
                     cudaStream_t st);

template <typename T, bool P_SMEM, int POLY>
static int launch_k(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    cudaStream_t st) {
    auto kern = xattn_fwd_kernel<T, P_SMEM, POLY>;
    const int smem = (P_SMEM ? XattnSmem::bytes_smem_variant : XattnSmem::bytes_tmem_variant) + 1024;
    static bool configured = false;      // per instantiation; cudaFuncSetAttribute is not a stream operation
    if (!configured) {
        RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    kern<<<grid, kThreads, smem, st>>>(mq, mk, mv, prm);
    RD_LAUNCH_CHECK("xattn_fwd_kernel");
    return RD_OK;
}

constexpr int kDefaultPoly = 1;   // polynomial path off: measured slower on B200 (attention log)

template <typename T>
static int launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    int flags, cudaStream_t st) {
    if (flags & RD_XATTN_P_VIA_SMEM) return launch_k<T, true, 0>(mq, mk, mv, prm, grid, st);
    int poly = (flags >> 8) & 0xF;       // RD_XATTN_POLY(n): 0 = default, 1 = off, n >= 2 = every n-th pair
    if (poly == 0) poly = kDefaultPoly;
    switch (poly) {
        case 1: return launch_k<T, false, 0>(mq, mk, mv, prm, grid, st);
        case 2: return launch_k<T, false, 2>(mq, mk, mv, prm, grid, st);
        case 3: return launch_k<T, false, 3>(mq, mk, mv, prm, grid, st);
        case 4: return launch_k<T, false, 4>(mq, mk, mv, prm, grid, st);
        case 6: return launch_k<T, false, 6>(mq, mk, mv, prm, grid, st);
        default: set_error("RD_XATTN_POLY(%d) is not instantiated (1, 2, 3, 4, 6)", poly); return RD_ERR_UNSUPPORTED;
    }
}

}  // namespace rd

using namespace rd;

extern "C" int rd_xattn_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                            int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t k_sb, int64_t k_ss, int64_t k_sh,
                            int64_t v_sb, int64_t v_ss, int64_t v_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh,
                            rd_dtype dtype, float scale, int flags, void* stream) {
    RD_CHECK_ARG(q && k && v && o, "NULL pointer");
    RD_CHECK_ARG(B >= 1 && H >= 1 && Sq >= 1 && Sk >= 1, "bad sizes B=%d H=%d Sq=%d Sk=%d", B, H, Sq, Sk);
    RD_CHECK_ARG(H <= 65535 && B <= 65535, "H and B must be <= 65535");
    if (D != kD) { set_error("head_dim %d is not instantiated (only 64)", D); return RD_ERR_UNSUPPORTED; }
    RD_CHECK_ARG(scale > 0.f && isfinite(scale), "scale must be positive and finite");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RD_F32) {
        const int64_t qs[3] = {q_sb, q_ss, q_sh}, ks[3] = {k_sb, k_ss, k_sh}, vs[3] = {v_sb, v_ss, v_sh}, os[3] = {o_sb, o_ss, o_sh};
        return xattn_f32_launch((const float*)q, (const float*)k, (const float*)v, (float*)o, B, H, Sq, Sk, qs, ks, vs, os,
                                scale, st);
    }
    RD_CHECK_ARG(dtype == RD_F16 || dtype == RD_BF16, "bad dtype %d", (int)dtype);
    // TMA / 16-byte vector store alignment
    const int64_t all[12] = {q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, v_sb, v_ss, v_sh, o_sb, o_ss, o_sh};
    for (int i = 0; i < 12; ++i) RD_CHECK_ARG(all[i] >= 0 && (all[i] % 8) == 0, "stride %d (=%lld) must be a non-negative multiple of 8 elements", i, (long long)all[i]);
    RD_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)o % 16) == 0,
                 "q/k/v/o must be 16-byte aligned");
    CUtensorMap mq, mk, mv;
    if (int e = make_qkv_map(&mq, q, dtype, B, H, Sq, q_sb, q_ss, q_sh, kBM)) return e;
    if (int e = make_qkv_map(&mk, k, dtype, B, H, Sk, k_sb, k_ss, k_sh, kBN)) return e;
    if (int e = make_qkv_map(&mv, v, dtype, B, H, Sk, v_sb, v_ss, v_sh, kBN)) return e;
    XattnParams prm{};
    prm.o = o; prm.o_sb = o_sb; prm.o_ss = o_ss; prm.o_sh = o_sh;
    prm.Sq = Sq; prm.Sk = Sk; prm.n_kv_tiles = (Sk + kBN - 1) / kBN;
    prm.scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((Sq + kBM - 1) / kBM, H, B);
    if (dtype == RD_F16) return launch_t<__half>(mq, mk, mv, prm, grid, flags, st);
    return launch_t<__nv_bfloat16>(mq, mk, mv, prm, grid, flags, st);
}
