// Cross-frame self-attention forward for sm_100a, THREE-query-tile organisation (see xattn_sm100.cu for the
// two-tile kernel, the common design and the reference call sites this replaces).
//
// One CTA per SM = three 128-row query tiles of one (batch, head) sharing every K/V tile; the softmax works in steps of
// 64 keys (half a 128-key TMA tile) so that three tiles fit TMEM: per tile S fp32 64 columns, O fp32 64, P packed 32
// (160 x 3 = 480 of 512 columns).  512 threads x 128 registers, no setmaxnreg:
//   warps 0-11   softmax, warpgroup t = query tile t (one query row = one TMEM lane per thread): THREE softmax warps per
//                SM sub-partition share its MUFU pipe, which hides the ~900 cycles of non-MUFU work per step (barrier
//                waits, TMEM load, row maximum, TMEM store) better than two: 80 % -> 85-88 % MUFU utilisation on
//                many-wave grids (+2-5 %), at the price of a coarser grid (384 query rows per CTA) that loses on
//                grids of a few waves.  rd_xattn_fwd picks the kernel by grid size.
//   warps 12-14  one tcgen05.mma issuer per tile (uniform control flow, elected lane issues)
//   warp 15      TMA producer (Q tiles once; K/V tiles of 128 keys = two steps through 3-stage rings)
#include <cuda.h>
#include <math.h>
#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.cuh"
#include "xattn_common.cuh"

namespace rd {
namespace q3 {

using namespace ptx;
using namespace xa;

#ifdef RD_XATTN_TRACE
__device__ unsigned long long rd_xattn_trace[16];   // see xattn_common.cuh
#endif

constexpr int kBM = 128;       // rows per query tile
constexpr int kNQ = 3;         // query tiles per CTA
constexpr int kBN = 128;       // keys per K/V tile (one TMA box)
constexpr int kBK = 64;        // keys per softmax / MMA step
constexpr int kStepBytes = kBK * 64 * 2;
constexpr int kD = 64;         // head dim
constexpr int kStages = 3;
constexpr int kTileBytes = kBN * kD * 2;   // 16384
constexpr int kSoftmaxThreads = kNQ * 128;
constexpr int kThreads = kSoftmaxThreads + 128;   // last warpgroup: three MMA issuer warps + the TMA warp
constexpr int kMmaWarp = kNQ * 4, kTmaWarp = kNQ * 4 + 3;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColTile = 160, kColS = 0, kColO = 64, kColP = 128;   // per tile: S 64, O 64, P 32 columns
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
    int Sq, Sk, n_kv_tiles, n_steps;
    float scale_log2;   // scale * log2(e)
};


// row maximum of one 64-key step: 8 independent chains of 3-input max (one dependent chain would be ~250 cycles)
__device__ __forceinline__ float row_max64(const uint32_t (&a)[32], const uint32_t (&b)[32]) {
    float mx[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) mx[q] = max3(__uint_as_float(a[4 * q]), __uint_as_float(a[4 * q + 1]), __uint_as_float(a[4 * q + 2]));
#pragma unroll
    for (int q = 0; q < 8; ++q) mx[q] = max3(mx[q], __uint_as_float(a[4 * q + 3]), __uint_as_float(b[4 * q]));
#pragma unroll
    for (int q = 0; q < 8; ++q) mx[q] = max3(mx[q], __uint_as_float(b[4 * q + 1]), __uint_as_float(b[4 * q + 2]));
#pragma unroll
    for (int q = 0; q < 8; ++q) mx[q] = fmaxf(mx[q], __uint_as_float(b[4 * q + 3]));
    return fmaxf(max3(mx[0], mx[1], mx[2]), max3(max3(mx[3], mx[4], mx[5]), mx[6], mx[7]));
}

__device__ __forceinline__ void mask_tail64(uint32_t (&a)[32], uint32_t (&b)[32], int valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i >= valid) a[i] = 0xff800000u;        // -inf
        if (32 + i >= valid) b[i] = 0xff800000u;
    }
}

// POLY: every POLY-th score pair takes the polynomial exp2 path (0 = all on MUFU)
template <typename T, int POLY>
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
    const int n_tiles = prm.n_kv_tiles, n_steps = prm.n_steps;

    if (threadIdx.x == 0) {
        mbar_init(&bars[Q_FULL], 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars[K_FULL + s], 1);
            mbar_init(&bars[K_EMPTY + s], kNQ);    // one tcgen05.commit per issuer
            mbar_init(&bars[V_FULL + s], 1);
            mbar_init(&bars[V_EMPTY + s], kNQ);
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
        } else if (warp < kMmaWarp + kNQ) {
        // ===================================================== MMA issuer of query tile t (uniform control flow)
        {
            const bool leader = elect_one();
            constexpr uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            constexpr uint32_t idesc_qk = make_idesc_f16(fmt, 0, 0, kBM, kBK);
            constexpr uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, kBM, kD);   // B = V is MN-major
            const int t = warp - kMmaWarp;
            const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0) + t * kColTile;
            const uint64_t dq = make_smem_desc_sw128(smem_u32(smem + XattnSmem::q + t * kTileBytes), 16, 1024);

            auto issue_qk = [&](int i) {
                const int tile = i >> 1, st = tile % kStages, b = i & 1;
                mbar_wait(&bars[K_FULL + st], (tile / kStages) & 1);
                if (i > 0) mbar_wait(&bars[S_FREE + t], (i - 1) & 1);   // softmax_t(i-1) has S in registers
                tc_fence_after();
                const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + XattnSmem::k + st * kTileBytes + b * kStepBytes), 16, 1024);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)             // +32 bytes along K inside the 128B swizzle row
                        mma_f16_ss(tm + kColS, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0);
                    if (b == 1 || i == n_steps - 1) mma_commit(&bars[K_EMPTY + st]);
                    mma_commit(&bars[S_FULL + t]);
                }
                __syncwarp();
            };

            mbar_wait(&bars[Q_FULL], 0);
            issue_qk(0);
            for (int j = 0; j < n_steps; ++j) {
                if (j + 1 < n_steps) issue_qk(j + 1);
                const int tile = j >> 1, st = tile % kStages, b = j & 1;
                mbar_wait(&bars[V_FULL + st], (tile / kStages) & 1);
                mbar_wait(&bars[P_FULL + t], j & 1);
                tc_fence_after();
                const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + XattnSmem::v + st * kTileBytes + b * kStepBytes), 16, 1024);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kBK / 16; ++kk)            // V advances 16 key rows = 2048 bytes per K step
                        mma_f16_ts(tm + kColO, tm + kColP + 8 * kk, dv + (uint64_t)(128 * kk), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                    if (b == 1 || j == n_steps - 1) mma_commit(&bars[V_EMPTY + st]);
                    mma_commit(&bars[O_READY + t]);
                }
                __syncwarp();
            }
        }
        }
    } else {
        // ===================================================== softmax warps (one query row per thread)
        const int t = warp >> 2;                       // query tile of this warpgroup
        const int row = (warp & 3) * 32 + lane;        // row inside the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + t * kColTile + kColS, tO = tmem_base + lane_off + t * kColTile + kColO,
                       tP = tmem_base + lane_off + t * kColTile + kColP;
        uint64_t* s_full = &bars[S_FULL + t];
        uint64_t* s_free = &bars[S_FREE + t];
        uint64_t* p_full = &bars[P_FULL + t];
        uint64_t* o_ready = &bars[O_READY + t];
        const float sl2 = prm.scale_log2;
        const uint64_t sl2_2 = pack2(sl2, sl2);
        float m_used = -INFINITY;   // raw-score units
        float l = 0.f;
        RD_TRACE_DECL

        for (int j = 0; j < n_steps; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            RD_TRACE(0)
            uint32_t sa[32], sb[32];
            tmem_ld32(tS + 0, sa);
            tmem_ld32(tS + 32, sb);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
            RD_TRACE(1)
            const int valid = prm.Sk - j * kBK;   // keys of this step that exist
            if (valid < kBK) mask_tail64(sa, sb, valid);
            const float m_new = fmaxf(m_used, row_max64(sa, sb));
            RD_TRACE(2)
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
            uint64_t ls2[2] = {0ull, 0ull};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const uint32_t r0 = h == 0 ? sa[i] : sb[i], r1 = h == 0 ? sa[i + 1] : sb[i + 1];
                    const uint64_t x2 = ffma2(pack2(__uint_as_float(r0), __uint_as_float(r1)), sl2_2, nmb2);
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
                    pk[i >> 1] = PackPair<T>::pack(p0, p1);
                }
                tmem_st16(tP + h * 16, pk);
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
            RD_TRACE(6)
        }
        RD_TRACE_FLUSH

        // ------------------------------------------------- epilogue: O / l -> global
        mbar_wait(o_ready, (n_steps - 1) & 1);
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
// host side (argument checks are done by rd_xattn_fwd in xattn_sm100.cu)
// ---------------------------------------------------------------------------------------------
template <typename T, int POLY>
static int launch_k(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    cudaStream_t st) {
    auto kern = xattn_fwd_kernel<T, POLY>;
    const int smem = XattnSmem::bytes + 1024;
    static PerDeviceOnce configured;      // cudaFuncSetAttribute is per device and not a stream operation
    if (configured.needed()) {
        RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.done();
    }
    kern<<<grid, kThreads, smem, st>>>(mq, mk, mv, prm);
    RD_LAUNCH_CHECK("xattn_fwd_kernel (three query tiles)");
    return RD_OK;
}

template <typename T>
static int launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    int flags, cudaStream_t st) {
    const int poly = (flags >> 8) & 0xF;       // RD_XATTN_POLY(n): 0 / 1 = off, 4 / 8 = every n-th pair
    switch (poly) {
        case 0: case 1: return launch_k<T, 0>(mq, mk, mv, prm, grid, st);
        case 4: return launch_k<T, 4>(mq, mk, mv, prm, grid, st);
        case 8: return launch_k<T, 8>(mq, mk, mv, prm, grid, st);
        default: set_error("RD_XATTN_POLY(%d) is not instantiated for the three-tile kernel (1, 4, 8)", poly); return RD_ERR_UNSUPPORTED;
    }
}

}  // namespace q3

int xattn_q3_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk,
                 int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t k_sb, int64_t k_ss, int64_t k_sh,
                 int64_t v_sb, int64_t v_ss, int64_t v_sh, int64_t o_sb, int64_t o_ss, int64_t o_sh,
                 rd_dtype dtype, float scale, int flags, cudaStream_t st) {
    using namespace q3;
    CUtensorMap mq, mk, mv;
    if (int e = make_qkv_map(&mq, q, dtype, B, H, Sq, q_sb, q_ss, q_sh, kBM)) return e;
    if (int e = make_qkv_map(&mk, k, dtype, B, H, Sk, k_sb, k_ss, k_sh, kBN)) return e;
    if (int e = make_qkv_map(&mv, v, dtype, B, H, Sk, v_sb, v_ss, v_sh, kBN)) return e;
    XattnParams prm{};
    prm.o = o; prm.o_sb = o_sb; prm.o_ss = o_ss; prm.o_sh = o_sh;
    prm.Sq = Sq; prm.Sk = Sk; prm.n_kv_tiles = (Sk + kBN - 1) / kBN; prm.n_steps = (Sk + kBK - 1) / kBK;
    prm.scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((Sq + kNQ * kBM - 1) / (kNQ * kBM), H, B);
    if (dtype == RD_F16) return launch_t<__half>(mq, mk, mv, prm, grid, flags, st);
    return launch_t<__nv_bfloat16>(mq, mk, mv, prm, grid, flags, st);
}

}  // namespace rd
