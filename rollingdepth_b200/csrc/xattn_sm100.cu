// Cross-frame self-attention forward for sm_100a: flash-style, head_dim 64, fp16/bf16 in,
// fp32 accumulate.  Replaces the fused-attention library call of the reference's processors
// (diffusers/src/diffusers/models/attention_processor.py:2251-2253 and :2036-2038).
//
// At head_dim 64 one 128x128 score tile costs 512 tensor-pipe cycles (QK^T + PV) but 1024 MUFU cycles
// (16 exponentials / clk / SM), so the kernel is organised around keeping the MUFU pipe busy:
//
// One CTA per SM = TWO 128-row query tiles of one (batch, head) that share every K/V tile.  Warp roles (384 threads;
// setmaxnreg moves registers from the third warpgroup to the softmax warpgroups: 88 vs 208 per thread):
//   warps 0-3  softmax of query tile 0, warps 4-7 softmax of query tile 1.  Each thread owns one query row
//              (= one TMEM lane): S -> registers (tcgen05.ld), row maximum, lazy (thresholded) rescaling of O
//              in TMEM, exponentials, P -> TMEM as packed fp16/bf16 (tcgen05.st), epilogue O / l -> global.
//              The two warpgroups run free (two warps per SM sub-partition share its MUFU pipe; while one is
//              in its TMEM loads / row maximum / barrier traffic the other one has the pipe to itself).  A forced
//              ping-pong on named barriers (RD_XATTN_PINGPONG) was measured slower and is kept as an A/B switch.
//   warps 8,9  tcgen05.mma issuers (one elected thread each; warp 8 also owns TMEM), ONE PER QUERY TILE so that the
//              two softmax warpgroups are decoupled (an in-order issuer shared by both tiles makes each warpgroup wait
//              for the other one's P): S_t = Q_t K^T (SS mode, both K-major, SWIZZLE_128B), O_t += P_t V (A = P_t
//              from TMEM, B = V MN-major from shared memory).  K/V stages are released when both issuers committed.
//   warp 10    TMA producer: both Q tiles once, then K/V tiles of 128 keys through 3-stage rings.
// TMEM (all 512 columns): S_t fp32 [128 t, +128), O_t fp32 [256 + 64 t, +64), P_t packed 16-bit [384 + 64 t, +64).
// QK^T of step j+1 is issued as soon as the softmax warps have pulled S_j into registers.
// Element-wise softmax arithmetic uses the packed fp32x2 instructions of sm_100 (FFMA2 / FADD2), and a tunable
// share of the exponentials is evaluated by a degree-3 polynomial on the FMA pipe instead of MUFU.EX2.
#include <cuda.h>
#include <math.h>
#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.cuh"
#include "xattn_common.cuh"

// internal A/B switches of rd_xattn_fwd `flags` (not part of the public header)
#define RD_XATTN_PINGPONG 1
#define RD_XATTN_SINGLE_ISSUER 2

namespace rd {

using namespace ptx;
using namespace xa;

#ifdef RD_XATTN_TRACE
__device__ unsigned long long rd_xattn_trace[16];   // see xattn_common.cuh
#endif

constexpr int kBM = 128;       // rows per query tile
constexpr int kNQ = 2;         // query tiles per CTA
constexpr int kBN = 128;       // keys per K/V tile
constexpr int kD = 64;         // head dim
constexpr int kStages = 3;
constexpr int kTileBytes = kBN * kD * 2;   // 16384
constexpr int kSoftmaxThreads = kNQ * 128;
constexpr int kThreads = kSoftmaxThreads + 128;   // third warpgroup: two MMA warps, TMA warp, one idle warp (register donors)
constexpr int kMmaWarp = kNQ * 4, kTmaWarp = kNQ * 4 + 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;   // + 128 t, + 64 t, + 64 t
constexpr float kRescaleThreshold = 8.0f;  // log2 units; p stays <= 2^8

struct XattnSmem {
    // offsets inside the 1024-byte aligned dynamic shared memory block
    static constexpr int q = 0;                                   // kNQ tiles
    static constexpr int k = q + kNQ * kTileBytes;
    static constexpr int v = k + kStages * kTileBytes;
    static constexpr int bars = v + kStages * kTileBytes;         // 32 x 8 bytes
    static constexpr int tmem_ptr = bars + 32 * 8;
    static constexpr int bytes = tmem_ptr + 16;
};
// barrier indices
constexpr int Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + kStages, V_FULL = K_EMPTY + kStages, V_EMPTY = V_FULL + kStages,
              S_FULL = V_EMPTY + kStages, S_FREE = S_FULL + kNQ, P_FULL = S_FREE + kNQ, O_READY = P_FULL + kNQ;
static_assert(O_READY + kNQ <= 32, "barrier table too small");

struct XattnParams {
    void* o;
    long long o_sb, o_ss, o_sh;
    int Sq, Sk, n_kv_tiles;
    float scale_log2;   // scale * log2(e)
};


__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// POLY: every POLY-th score pair takes the polynomial exp2 path (0 = all on MUFU)
// PINGPONG: the two softmax warpgroups alternate their exponential phases (named barriers 1 and 2)
// SPLIT: one MMA issuer warp per query tile (false: a single in-order issuer for both tiles)
template <typename T, int POLY, bool PINGPONG, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
xattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const XattnParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XattnSmem::bars);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + XattnSmem::tmem_ptr);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_tiles = prm.n_kv_tiles;

    if (threadIdx.x == 0) {
        mbar_init(&bars[Q_FULL], 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars[K_FULL + s], 1);
            mbar_init(&bars[K_EMPTY + s], SPLIT ? kNQ : 1);    // one tcgen05.commit per issuer
            mbar_init(&bars[V_FULL + s], 1);
            mbar_init(&bars[V_EMPTY + s], SPLIT ? kNQ : 1);
        }
        for (int t = 0; t < kNQ; ++t) {
            mbar_init(&bars[S_FULL + t], 1);
            mbar_init(&bars[S_FREE + t], 4);     // one elected lane per softmax warp
            mbar_init(&bars[P_FULL + t], 4);
            mbar_init(&bars[O_READY + t], 1);
        }
        mbar_fence_init();
    }
    if (warp == kTmaWarp && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
    }
    if (warp == kMmaWarp) {
        tmem_alloc(tmem_ptr_smem, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // (setmaxnreg must dominate the code it applies to: one per role branch, no re-convergence in between)
    if (warp >= kMmaWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        if (warp == kTmaWarp) {
        // ===================================================== TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars[Q_FULL], kNQ * kTileBytes);
            for (int t = 0; t < kNQ; ++t)
                tma_load_4d(smem + XattnSmem::q + t * kTileBytes, &tm_q, &bars[Q_FULL], 0, (q_blk * kNQ + t) * kBM, head, batch);
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
        } else if (warp == kMmaWarp || (SPLIT && warp == kMmaWarp + 1)) {
        // ===================================================== MMA issuer(s): query tiles [t_lo, t_hi)
        // All 32 lanes run the loop (uniform control flow, descriptors in uniform registers); one elected lane issues.
        {
            const bool leader = elect_one();
            constexpr uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            constexpr uint32_t idesc_qk = make_idesc_f16(fmt, 0, 0, kBM, kBN);
            constexpr uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, kBM, kD);   // B = V is MN-major
            const int t_lo = SPLIT ? warp - kMmaWarp : 0, t_hi = SPLIT ? t_lo + 1 : kNQ;
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);

            auto issue_qk = [&](int j) {
                const int st = j % kStages;
                mbar_wait(&bars[K_FULL + st], (j / kStages) & 1);
                const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + XattnSmem::k + st * kTileBytes), 16, 1024);
                for (int t = t_lo; t < t_hi; ++t) {
                    if (j > 0) mbar_wait(&bars[S_FREE + t], (j - 1) & 1);   // softmax_t(j-1) has S in registers
                    tc_fence_after();
                    const uint64_t dq = make_smem_desc_sw128(smem_u32(smem + XattnSmem::q + t * kTileBytes), 16, 1024);
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < kD / 16; ++kk)             // +32 bytes along K inside the 128B swizzle row
                            mma_f16_ss(tm + kColS + t * 128, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0);
                        if (t == t_hi - 1) mma_commit(&bars[K_EMPTY + st]);
                        mma_commit(&bars[S_FULL + t]);
                    }
                    __syncwarp();
                }
            };

            mbar_wait(&bars[Q_FULL], 0);
            issue_qk(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_qk(j + 1);
                const int st = j % kStages;
                mbar_wait(&bars[V_FULL + st], (j / kStages) & 1);
                const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + XattnSmem::v + st * kTileBytes), 16, 1024);
                for (int t = t_lo; t < t_hi; ++t) {
                    mbar_wait(&bars[P_FULL + t], j & 1);
                    tc_fence_after();
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < kBN / 16; ++kk)            // V advances 16 key rows = 2048 bytes per K step
                            mma_f16_ts(tm + kColO + t * 64, tm + kColP + t * 64 + 8 * kk, dv + (uint64_t)(128 * kk),
                                       idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                        if (t == t_hi - 1) mma_commit(&bars[V_EMPTY + st]);
                        mma_commit(&bars[O_READY + t]);
                    }
                    __syncwarp();
                }
            }
        }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ===================================================== softmax warps (one query row per thread)
        const int t = warp >> 2;                       // query tile of this warpgroup
        const int row = (warp & 3) * 32 + lane;        // row inside the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + kColS + t * 128, tO = tmem_base + lane_off + kColO + t * 64,
                       tP = tmem_base + lane_off + kColP + t * 64;
        uint64_t* s_full = &bars[S_FULL + t];
        uint64_t* s_free = &bars[S_FREE + t];
        uint64_t* p_full = &bars[P_FULL + t];
        uint64_t* o_ready = &bars[O_READY + t];
        const float sl2 = prm.scale_log2;
        const uint64_t sl2_2 = pack2(sl2, sl2);
        float m_used = -INFINITY;   // raw-score units
        float l = 0.f;
        RD_TRACE_DECL
        if (PINGPONG && t == 1) named_bar_arrive(1, kSoftmaxThreads);   // warpgroup 0 goes first

        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            RD_TRACE(0)
            uint32_t sr[4][32];
            tmem_ld32(tS + 0, sr[0]);
            tmem_ld32(tS + 32, sr[1]);
            tmem_ld32(tS + 64, sr[2]);
            tmem_ld32(tS + 96, sr[3]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
            RD_TRACE(1)

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
            RD_TRACE(2)

            if (PINGPONG) named_bar_sync(1 + t, kSoftmaxThreads);      // our turn on the MUFU pipe
            if (j > 0) {
                // P (and O) may only be touched once PV(j-1) has completed
                mbar_wait(o_ready, (j - 1) & 1);
                tc_fence_after();
                RD_TRACE(3)
                const bool need = (m_new - m_used) * sl2 > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    const float f = ex2((m_used - m_new) * sl2);
                    l *= f;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {           // rare path: keep its register footprint small
                        uint32_t orr[32];
                        tmem_ld32(tO + c * 32, orr);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
                        tmem_st32(tO + c * 32, orr);
                    }
                    tmem_st_wait();
                    m_used = m_new;
                }
            } else {
                m_used = m_new;
            }
            RD_TRACE(4)

            const float mb = m_used * sl2;
            const uint64_t nmb2 = pack2(-mb, -mb);
            uint64_t ls2[2] = {0ull, 0ull};            // independent partial row sums (two fp32 each)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t pk[32];
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = half * 2 + cc;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const uint64_t x2 = ffma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), sl2_2, nmb2);
                        float p0, p1;
                        if (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == POLY - 1) {
                            ex2_poly2(x2, p0, p1);
                        } else {
                            float x0, x1;
                            unpack2(x2, x0, x1);
                            p0 = ex2(x0);
                            p1 = ex2(x1);
                        }
                        ls2[(i >> 1) & 1] = fadd2(ls2[(i >> 1) & 1], pack2(p0, p1));
                        pk[cc * 16 + (i >> 1)] = PackPair<T>::pack(p0, p1);
                    }
                }
                tmem_st32(tP + half * 32, pk);
            }
            {
                float a0, a1, b0, b1;
                unpack2(ls2[0], a0, a1);
                unpack2(ls2[1], b0, b1);
                l += (a0 + a1) + (b0 + b1);
            }
            RD_TRACE(5)
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            if (PINGPONG && !(t == 1 && j == n_tiles - 1)) named_bar_arrive(1 + (t ^ 1), kSoftmaxThreads);
            RD_TRACE(6)
        }
        RD_TRACE_FLUSH

        // ------------------------------------------------- epilogue: O / l -> global
        mbar_wait(o_ready, (n_tiles - 1) & 1);
        tc_fence_after();
        uint32_t orr[2][32];
        tmem_ld32(tO + 0, orr[0]);
        tmem_ld32(tO + 32, orr[1]);
        tmem_ld_wait();
        const float inv_l = 1.0f / l;
        const int grow = (q_blk * kNQ + t) * kBM + row;
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
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side (tensor maps: tensormap.cuh)
// ---------------------------------------------------------------------------------------------
// three-query-tile kernel (xattn_sm100_q3.cu)
int xattn_q3_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk,
                 int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t k_sb, int64_t k_ss, int64_t k_sh,
                 int64_t v_sb, int64_t v_ss, int64_t v_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh,
                 rd_dtype dtype, float scale, int flags, cudaStream_t st);

int xattn_f32_launch(const float* q, const float* k, const float* v, float* o, int B, int H, int Sq, int Sk,
                     const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, float scale,
                     cudaStream_t st);
int xattn_tf32_launch(const float* q, const float* k, const float* v, float* o, int B, int H, int Sq, int Sk,
                      const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, float scale,
                      cudaStream_t st);

template <typename T, int POLY, bool PINGPONG, bool SPLIT>
static int launch_k(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    cudaStream_t st) {
    auto kern = xattn_fwd_kernel<T, POLY, PINGPONG, SPLIT>;
    const int smem = XattnSmem::bytes + 1024;
    static PerDeviceOnce configured;      // cudaFuncSetAttribute is per device and not a stream operation
    if (configured.needed()) {
        RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.done();
    }
    kern<<<grid, kThreads, smem, st>>>(mq, mk, mv, prm);
    RD_LAUNCH_CHECK("xattn_fwd_kernel");
    return RD_OK;
}

constexpr int kDefaultPoly = 1;   // polynomial path off by default (attention log)
constexpr bool kDefaultF32Tensor = true;    // fp32 inputs: the 3xTF32 tcgen05 kernel (xattn_tf32.cu); RD_XATTN_F32(1) = scalar FFMA kernel

template <typename T>
static int launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    int flags, cudaStream_t st) {
    int poly = (flags >> 8) & 0xF;       // RD_XATTN_POLY(n): 0 = default, 1 = off, n >= 2 = every n-th pair
    if (poly == 0) poly = kDefaultPoly;
    if (flags & RD_XATTN_PINGPONG) {
        if (poly == 4) return launch_k<T, 4, true, true>(mq, mk, mv, prm, grid, st);
        return launch_k<T, 0, true, true>(mq, mk, mv, prm, grid, st);
    }
    if (flags & RD_XATTN_SINGLE_ISSUER) return launch_k<T, 0, false, false>(mq, mk, mv, prm, grid, st);
    switch (poly) {
        case 1: return launch_k<T, 0, false, true>(mq, mk, mv, prm, grid, st);
        case 2: return launch_k<T, 2, false, true>(mq, mk, mv, prm, grid, st);
        case 3: return launch_k<T, 3, false, true>(mq, mk, mv, prm, grid, st);
        case 4: return launch_k<T, 4, false, true>(mq, mk, mv, prm, grid, st);
        case 6: return launch_k<T, 6, false, true>(mq, mk, mv, prm, grid, st);
        case 8: return launch_k<T, 8, false, true>(mq, mk, mv, prm, grid, st);
        default: set_error("RD_XATTN_POLY(%d) is not instantiated (1, 2, 3, 4, 6, 8)", poly); return RD_ERR_UNSUPPORTED;
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
        // RD_XATTN_F32(1): the scalar FFMA kernel, (2): the 3xTF32 tcgen05 kernel, 0: default
        const int f32_path = (flags >> 14) & 3;
        if (f32_path == 2 || (f32_path == 0 && kDefaultF32Tensor))
            return xattn_tf32_launch((const float*)q, (const float*)k, (const float*)v, (float*)o, B, H, Sq, Sk, qs, ks, vs, os,
                                     scale, st);
        return xattn_f32_launch((const float*)q, (const float*)k, (const float*)v, (float*)o, B, H, Sq, Sk, qs, ks, vs, os,
                                scale, st);
    }
    RD_CHECK_ARG(dtype == RD_F16 || dtype == RD_BF16, "bad dtype %d", (int)dtype);
    // TMA / 16-byte vector store alignment
    const int64_t all[12] = {q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, v_sb, v_ss, v_sh, o_sb, o_ss, o_sh};
    for (int i = 0; i < 12; ++i) RD_CHECK_ARG(all[i] >= 0 && (all[i] % 8) == 0, "stride %d (=%lld) must be a non-negative multiple of 8 elements", i, (long long)all[i]);
    RD_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)o % 16) == 0,
                 "q/k/v/o must be 16-byte aligned");
    {   // kernel choice: three query tiles per CTA (384 rows) pay off on many-wave grids of long sequences; two tiles
        // (256 rows) quantise better on small grids.  RD_XATTN_TILES(2 / 3) forces one (A/B, parity tests of both).
        const int tiles = (flags >> 4) & 3;
        const long long ctas3 = (long long)((Sq + 383) / 384) * H * B;
        const bool big = ctas3 >= 4ll * num_sms() && Sk >= 2048;
        const int poly = (flags >> 8) & 0xF;
        if ((tiles == 3 || (tiles == 0 && big)) && !(flags & 3) && (poly == 0 || poly == 1 || poly == 4 || poly == 8))
            return xattn_q3_fwd(q, k, v, o, B, H, Sq, Sk, q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, v_sb, v_ss, v_sh, o_sb, o_ss, o_sh,
                                dtype, scale, flags, st);
    }
    CUtensorMap mq, mk, mv;
    if (int e = make_qkv_map(&mq, q, dtype, B, H, Sq, q_sb, q_ss, q_sh, kBM)) return e;
    if (int e = make_qkv_map(&mk, k, dtype, B, H, Sk, k_sb, k_ss, k_sh, kBN)) return e;
    if (int e = make_qkv_map(&mv, v, dtype, B, H, Sk, v_sb, v_ss, v_sh, kBN)) return e;
    XattnParams prm{};
    prm.o = o; prm.o_sb = o_sb; prm.o_ss = o_ss; prm.o_sh = o_sh;
    prm.Sq = Sq; prm.Sk = Sk; prm.n_kv_tiles = (Sk + kBN - 1) / kBN;
    prm.scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((Sq + kNQ * kBM - 1) / (kNQ * kBM), H, B);
    if (dtype == RD_F16) return launch_t<__half>(mq, mk, mv, prm, grid, flags, st);
    return launch_t<__nv_bfloat16>(mq, mk, mv, prm, grid, flags, st);
}
