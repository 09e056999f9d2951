// fp32 cross-frame attention on the tensor cores: error-compensated 3xTF32 (the `paper` preset and cfg1 run in fp32).
//
// A single TF32 pass (10-bit mantissa) is not accurate enough for the 1e-4 max-abs bound against the reference's fp32
// SDPA path (attention_processor.py:2251-2253; logit error ~1e-3 at head_dim 64).  Every fp32
// operand is therefore split into hi = the TF32-representable top 19 bits and lo = x - hi (exact), and each product is
// evaluated as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM (the dropped lo*lo term is ~2^-22 relative):
//     S  = Qhi Khi^T + Qhi Klo^T + Qlo Khi^T          O += Phi Vhi + Phi Vlo + Plo Vhi
// Same flash-style organisation as the fp16 kernels (xattn_sm100.cu), one 128-row query tile per CTA, 64-key steps:
//   warps 0-3   softmax: one query row = one TMEM lane per thread; P is split into hi / lo in registers and written to
//               TMEM (tcgen05.st) as the A operand of the PV products
//   warps 4-7   operand splitters: the raw fp32 tiles land in shared memory by TMA; these warps rewrite Q and K in place as
//               hi and write lo beside them (element-wise, so the SWIZZLE_128B layout is preserved), and TRANSPOSE V into
//               K-major [channel][key] hi / lo tiles (kind::tf32 takes no MN-major B operand: measured, the MMA is a no-op),
//               then fence.proxy.async + mbarrier so that tcgen05.mma may read them
//   warp 8      tcgen05.mma issuer (kind::tf32, cta_group::1; owns TMEM): 24 MMAs per step for S, 24 for O
//   warp 9      TMA producer: Q once, K/V tiles of 64 keys through a 2-stage ring (fp32 rows of 64 = two 128-byte
//               swizzle atoms of 32 channels each: two boxes per tile)
// TMEM: S fp32 64 columns, O 64, Phi 64, Plo 64.  Shared memory: Q hi/lo 64 KB + 2 x (K hi/lo + V hi/lo) 128 KB.
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.cuh"
#include "xattn_common.cuh"

namespace rd {
namespace tf32 {

using namespace ptx;
using namespace xa;

constexpr int kBM = 128;       // rows per query tile
constexpr int kBK = 64;        // keys per step
constexpr int kD = 64;         // head dim
constexpr int kStages = 2;
constexpr int kAtomQ = kBM * 128;          // bytes of one 32-channel swizzle atom of the Q tile
constexpr int kAtomK = kBK * 128;          // ... of a K / V tile
constexpr int kQBytes = 2 * kAtomQ, kKBytes = 2 * kAtomK;
constexpr int kThreads = 320;
constexpr int kSplitWarp0 = 4, kMmaWarp = 8, kTmaWarp = 9;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColO = 64, kColPhi = 128, kColPlo = 192;
constexpr float kRescaleThreshold = 8.0f;  // log2 units; p stays <= 2^8
constexpr uint32_t kHiMask = 0xFFFFE000u;  // sign + exponent + 10 mantissa bits = TF32

struct Smem {
    static constexpr int q_hi = 0;
    static constexpr int q_lo = q_hi + kQBytes;
    static constexpr int stage0 = q_lo + kQBytes;
    static constexpr int k_hi = 0, k_lo = kKBytes, v_hi = 2 * kKBytes, v_lo = 3 * kKBytes;   // offsets inside a stage
    static constexpr int stage_bytes = 4 * kKBytes;
    static constexpr int bars = stage0 + kStages * stage_bytes;       // 24 x 8 bytes
    static constexpr int tmem_ptr = bars + 24 * 8;
    static constexpr int bytes = tmem_ptr + 16;
};
// K and V travel through separate 2-stage rings: a K stage is free again as soon as QK^T of its step has completed, a V
// stage only after PV of its step -- with one ring for both the reload of step j+2 waited for PV(j) and every step paid the
// TMA + split latency in series (measured: 100 TFLOP/s; the MMA work of a step is a third of that)
constexpr int Q_RAW = 0, Q_SPLIT = 1, K_RAW = 2, K_SPLIT = K_RAW + kStages, K_EMPTY = K_SPLIT + kStages,
              V_RAW = K_EMPTY + kStages, V_SPLIT = V_RAW + kStages, V_EMPTY = V_SPLIT + kStages,
              S_FULL = V_EMPTY + kStages, S_FREE = S_FULL + 1, P_FULL = S_FREE + 1, O_READY = P_FULL + 1;
static_assert(O_READY < 24, "barrier table too small");

struct Params {
    float* o;
    long long o_sb, o_ss, o_sh;
    int Sq, Sk, n_steps;
    float scale_log2;
};

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor of kind::tf32: same bit layout as kind::f16, operand format code 2 = TF32
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t b_mn_major, uint32_t M, uint32_t N) {
    return make_idesc_f16(2u, 0u, b_mn_major, M, N);
}

// rewrite `n16` 16-byte chunks at `hi` in place as their TF32 part and write the remainders to `lo` (128 threads)
__device__ __forceinline__ void split_region(uint8_t* hi, uint8_t* lo, int n16, int tid) {
    uint4* h4 = reinterpret_cast<uint4*>(hi);
    uint4* l4 = reinterpret_cast<uint4*>(lo);
    for (int i = tid; i < n16; i += 128) {
        const uint4 r = h4[i];
        uint4 h, l;
        h.x = r.x & kHiMask; h.y = r.y & kHiMask; h.z = r.z & kHiMask; h.w = r.w & kHiMask;
        l.x = __float_as_uint(__uint_as_float(r.x) - __uint_as_float(h.x));
        l.y = __float_as_uint(__uint_as_float(r.y) - __uint_as_float(h.y));
        l.z = __float_as_uint(__uint_as_float(r.z) - __uint_as_float(h.z));
        l.w = __float_as_uint(__uint_as_float(r.w) - __uint_as_float(h.w));
        h4[i] = h;
        l4[i] = l;
    }
}

// V tile: raw [64 keys][64 channels] as TMA left it at `v` (two SWIZZLE_128B atoms of 32 channels, 128-byte rows = keys)
// -> V^T hi at `v` and V^T lo at `lo`, each [64 channels][64 keys] as two atoms of 32 keys (128-byte rows = channels): the
// K-major B operand of O += P V.  Thread (warp w, lane): channel 32 (w & 1) + lane, keys 32 (w >> 1) .. + 31; every thread
// pulls its 32 raw values into registers, the 128 splitter threads meet on a named barrier (the raw tile is overwritten in
// place), then each writes eight 16-byte chunks of hi and of lo.  Swizzle: 16-byte chunk index ^= (row & 7).
__device__ __forceinline__ void transpose_split_v(uint8_t* v, uint8_t* lo, int tid) {
    const int lane = tid & 31, w = tid >> 5;
    const int ch = 32 * (w & 1) + lane, a_in = w & 1, a_out = w >> 1;
    const uint8_t* in = v + a_in * kAtomK + (lane & 3) * 4;
    const int chunk_in = lane >> 2;
    uint32_t raw[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int key = 32 * a_out + i;
        raw[i] = *reinterpret_cast<const uint32_t*>(in + key * 128 + ((chunk_in ^ (key & 7)) << 4));
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    uint8_t* oh = v + a_out * kAtomK + ch * 128;
    uint8_t* ol = lo + a_out * kAtomK + ch * 128;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        uint4 h, l;
        h.x = raw[4 * c4] & kHiMask; h.y = raw[4 * c4 + 1] & kHiMask; h.z = raw[4 * c4 + 2] & kHiMask; h.w = raw[4 * c4 + 3] & kHiMask;
        l.x = __float_as_uint(__uint_as_float(raw[4 * c4]) - __uint_as_float(h.x));
        l.y = __float_as_uint(__uint_as_float(raw[4 * c4 + 1]) - __uint_as_float(h.y));
        l.z = __float_as_uint(__uint_as_float(raw[4 * c4 + 2]) - __uint_as_float(h.z));
        l.w = __float_as_uint(__uint_as_float(raw[4 * c4 + 3]) - __uint_as_float(h.w));
        const int off = (c4 ^ (ch & 7)) << 4;
        *reinterpret_cast<uint4*>(oh + off) = h;
        *reinterpret_cast<uint4*>(ol + off) = l;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
xattn_tf32_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + Smem::tmem_ptr);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_steps = prm.n_steps;

    if (threadIdx.x == 0) {
        mbar_init(&bars[Q_RAW], 1);
        mbar_init(&bars[Q_SPLIT], 4);            // one elected lane per splitter warp
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars[K_RAW + s], 1);
            mbar_init(&bars[K_SPLIT + s], 4);
            mbar_init(&bars[K_EMPTY + s], 1);    // tcgen05.commit after QK^T of the step
            mbar_init(&bars[V_RAW + s], 1);
            mbar_init(&bars[V_SPLIT + s], 4);
            mbar_init(&bars[V_EMPTY + s], 1);    // tcgen05.commit after PV of the step
        }
        mbar_init(&bars[S_FULL], 1);
        mbar_init(&bars[S_FREE], 4);             // one elected lane per softmax warp
        mbar_init(&bars[P_FULL], 4);
        mbar_init(&bars[O_READY], 1);
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

    if (warp == kTmaWarp) {
        // ===================================================== TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars[Q_RAW], kQBytes);
            tma_load_4d(smem + Smem::q_hi, &tm_q, &bars[Q_RAW], 0, q_blk * kBM, head, batch);
            tma_load_4d(smem + Smem::q_hi + kAtomQ, &tm_q, &bars[Q_RAW], 32, q_blk * kBM, head, batch);
            for (int j = 0; j < n_steps; ++j) {
                const int st = j % kStages;
                uint8_t* sb = smem + Smem::stage0 + st * Smem::stage_bytes;
                const uint32_t ph = ((j / kStages) & 1) ^ 1;
                mbar_wait(&bars[K_EMPTY + st], ph);
                mbar_arrive_expect_tx(&bars[K_RAW + st], kKBytes);
                tma_load_4d(sb + Smem::k_hi, &tm_k, &bars[K_RAW + st], 0, j * kBK, head, batch);
                tma_load_4d(sb + Smem::k_hi + kAtomK, &tm_k, &bars[K_RAW + st], 32, j * kBK, head, batch);
                mbar_wait(&bars[V_EMPTY + st], ph);
                mbar_arrive_expect_tx(&bars[V_RAW + st], kKBytes);
                tma_load_4d(sb + Smem::v_hi, &tm_v, &bars[V_RAW + st], 0, j * kBK, head, batch);
                tma_load_4d(sb + Smem::v_hi + kAtomK, &tm_v, &bars[V_RAW + st], 32, j * kBK, head, batch);
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================================================== MMA issuer (uniform control flow, elected lane issues)
        const bool leader = elect_one();
        constexpr uint32_t idesc_qk = make_idesc_tf32(0, kBM, kBK);      // S[128, 64 keys] = Q K^T, both K-major
        constexpr uint32_t idesc_pv = make_idesc_tf32(0, kBM, kD);       // O[128, 64 channels] += P V^T^T, B = V^T K-major
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);

        auto issue_qk = [&](int j) {
            const int st = j % kStages;
            uint8_t* sb = smem + Smem::stage0 + st * Smem::stage_bytes;
            mbar_wait(&bars[K_SPLIT + st], (j / kStages) & 1);
            if (j > 0) mbar_wait(&bars[S_FREE], (j - 1) & 1);            // softmax(j-1) has S in registers
            tc_fence_after();
            if (leader) {
                uint32_t acc = 0;
#pragma unroll
                for (int term = 0; term < 3; ++term) {                   // hi*hi, hi*lo, lo*hi
                    const uint8_t* qb = smem + (term == 2 ? Smem::q_lo : Smem::q_hi);
                    const uint8_t* kb = sb + (term == 1 ? Smem::k_lo : Smem::k_hi);
#pragma unroll
                    for (int a = 0; a < 2; ++a) {                        // the two 32-channel swizzle atoms of head_dim 64
                        const uint64_t dq = make_smem_desc_sw128(smem_u32(qb + a * kAtomQ), 16, 1024);
                        const uint64_t dk = make_smem_desc_sw128(smem_u32(kb + a * kAtomK), 16, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {                 // 8 channels = 32 bytes per MMA
                            mma_tf32_ss(tm + kColS, dq + 2 * kk, dk + 2 * kk, idesc_qk, acc);
                            acc = 1;
                        }
                    }
                }
                mma_commit(&bars[K_EMPTY + st]);
                mma_commit(&bars[S_FULL]);
            }
            __syncwarp();
        };

        mbar_wait(&bars[Q_SPLIT], 0);
        issue_qk(0);
        for (int j = 0; j < n_steps; ++j) {
            if (j + 1 < n_steps) issue_qk(j + 1);
            const int st = j % kStages;
            uint8_t* sb = smem + Smem::stage0 + st * Smem::stage_bytes;
            mbar_wait(&bars[V_SPLIT + st], (j / kStages) & 1);
            mbar_wait(&bars[P_FULL], j & 1);
            tc_fence_after();
            if (leader) {
                uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                for (int term = 0; term < 3; ++term) {                   // hi*hi, hi*lo, lo*hi
                    const uint32_t pa = tm + (term == 2 ? kColPlo : kColPhi);
                    const uint8_t* vb = sb + (term == 1 ? Smem::v_lo : Smem::v_hi);
#pragma unroll
                    for (int a = 0; a < 2; ++a) {                        // the two 32-key swizzle atoms of the transposed V tile
                        const uint64_t dv = make_smem_desc_sw128(smem_u32(vb + a * kAtomK), 16, 1024);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {                 // 8 keys = 32 bytes per MMA
                            mma_tf32_ts(tm + kColO, pa + a * 32 + 8 * kk, dv + 2 * kk, idesc_pv, acc);
                            acc = 1;
                        }
                    }
                }
                mma_commit(&bars[V_EMPTY + st]);
                mma_commit(&bars[O_READY]);
            }
            __syncwarp();
        }
    } else if (warp >= kSplitWarp0) {
        // ===================================================== operand splitters (128 threads)
        const int tid = threadIdx.x - kSplitWarp0 * 32;
        mbar_wait(&bars[Q_RAW], 0);
        split_region(smem + Smem::q_hi, smem + Smem::q_lo, kQBytes / 16, tid);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[Q_SPLIT]);
        for (int j = 0; j < n_steps; ++j) {
            const int st = j % kStages;
            uint8_t* sb = smem + Smem::stage0 + st * Smem::stage_bytes;
            mbar_wait(&bars[K_RAW + st], (j / kStages) & 1);
            split_region(sb + Smem::k_hi, sb + Smem::k_lo, kKBytes / 16, tid);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[K_SPLIT + st]);
            mbar_wait(&bars[V_RAW + st], (j / kStages) & 1);
            transpose_split_v(sb + Smem::v_hi, sb + Smem::v_lo, tid);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[V_SPLIT + st]);
        }
    } else {
        // ===================================================== softmax warps (one query row per thread)
        const int row = warp * 32 + lane;              // row inside the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + kColS, tO = tmem_base + lane_off + kColO,
                       tPhi = tmem_base + lane_off + kColPhi, tPlo = tmem_base + lane_off + kColPlo;
        const float sl2 = prm.scale_log2;
        float m_used = -INFINITY;   // raw-score units
        float l = 0.f;

        for (int j = 0; j < n_steps; ++j) {
            mbar_wait(&bars[S_FULL], j & 1);
            tc_fence_after();
            uint32_t sa[32], sb[32];
            tmem_ld32(tS + 0, sa);
            tmem_ld32(tS + 32, sb);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[S_FREE]);
            const int valid = prm.Sk - j * kBK;   // keys of this step that exist
            if (valid < kBK) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= valid) sa[i] = 0xff800000u;        // -inf
                    if (32 + i >= valid) sb[i] = 0xff800000u;
                }
            }
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                mx[(i >> 1) & 3] = max3(mx[(i >> 1) & 3], __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
                mx[(i >> 1) & 3] = max3(mx[(i >> 1) & 3], __uint_as_float(sb[i]), __uint_as_float(sb[i + 1]));
            }
            const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
            if (j > 0) {
                // P (and O) may only be touched once PV(j-1) has completed
                mbar_wait(&bars[O_READY], (j - 1) & 1);
                tc_fence_after();
                const bool need = (m_new - m_used) * sl2 > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    const float f = ex2((m_used - m_new) * sl2);
                    l *= f;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
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
            const float nmb = -m_used * sl2;
            float ls[2] = {0.f, 0.f};
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float p = ex2(fmaf(__uint_as_float(half == 0 ? sa[i] : sb[i]), sl2, nmb));
                    ls[i & 1] += p;
                    hi[i] = __float_as_uint(p) & kHiMask;
                    lo[i] = __float_as_uint(p - __uint_as_float(hi[i]));
                }
                tmem_st32(tPhi + half * 32, hi);
                tmem_st32(tPlo + half * 32, lo);
            }
            l += ls[0] + ls[1];
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[P_FULL]);
        }

        // ------------------------------------------------- epilogue: O / l -> global
        mbar_wait(&bars[O_READY], (n_steps - 1) & 1);
        tc_fence_after();
        uint32_t orr[2][32];
        tmem_ld32(tO + 0, orr[0]);
        tmem_ld32(tO + 32, orr[1]);
        tmem_ld_wait();
        const float inv_l = 1.0f / l;
        const int grow = q_blk * kBM + row;
        if (grow < prm.Sq) {
            float* dst = prm.o + (long long)batch * prm.o_sb + (long long)grow * prm.o_ss + (long long)head * prm.o_sh;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int e = c * 4;
                *reinterpret_cast<float4*>(dst + e) =
                    make_float4(__uint_as_float(orr[e >> 5][e & 31]) * inv_l, __uint_as_float(orr[e >> 5][(e & 31) + 1]) * inv_l,
                                __uint_as_float(orr[e >> 5][(e & 31) + 2]) * inv_l, __uint_as_float(orr[e >> 5][(e & 31) + 3]) * inv_l);
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

// 4-D map over a [B, S, H, 64] fp32 tensor: dims {D, S, H, B}, box {32, rows, 1, 1} (one 128-byte swizzle row per token
// and 32-channel half), OOB rows read as zero
static int make_f32_map(CUtensorMap* m, const void* base, int B, int H, int S, int64_t sb, int64_t ss, int64_t sh, int box_rows) {
    PFN_encodeTiled enc = tensormap_encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return RD_ERR_CUDA; }
    const cuuint64_t dims[4] = {64, (cuuint64_t)S, (cuuint64_t)H, (cuuint64_t)B};
    if (B == 1) sb = (int64_t)S * ss;   // unused stride; keep it legal
    if (H == 1) sh = 64;
    const cuuint64_t strides[3] = {(cuuint64_t)ss * 4, (cuuint64_t)sh * 4, (cuuint64_t)sb * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r); return RD_ERR_CUDA; }
    return RD_OK;
}

}  // namespace tf32

int xattn_tf32_launch(const float* q, const float* k, const float* v, float* o, int B, int H, int Sq, int Sk,
                      const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, float scale, cudaStream_t st) {
    using namespace tf32;
    for (int i = 0; i < 3; ++i)
        RD_CHECK_ARG(qs[i] % 4 == 0 && ks[i] % 4 == 0 && vs[i] % 4 == 0 && os[i] % 4 == 0,
                     "fp32 strides must be multiples of 4 elements");
    RD_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)o % 16) == 0,
                 "q/k/v/o must be 16-byte aligned");
    CUtensorMap mq, mk, mv;
    if (int e = make_f32_map(&mq, q, B, H, Sq, qs[0], qs[1], qs[2], kBM)) return e;
    if (int e = make_f32_map(&mk, k, B, H, Sk, ks[0], ks[1], ks[2], kBK)) return e;
    if (int e = make_f32_map(&mv, v, B, H, Sk, vs[0], vs[1], vs[2], kBK)) return e;
    Params prm{};
    prm.o = o; prm.o_sb = os[0]; prm.o_ss = os[1]; prm.o_sh = os[2];
    prm.Sq = Sq; prm.Sk = Sk; prm.n_steps = (Sk + kBK - 1) / kBK;
    prm.scale_log2 = scale * 1.4426950408889634f;
    const int smem = Smem::bytes + 1024;
    static PerDeviceOnce configured;      // cudaFuncSetAttribute is per device and not a stream operation
    if (configured.needed()) {
        RD_CUDA(cudaFuncSetAttribute(xattn_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.done();
    }
    dim3 grid((Sq + kBM - 1) / kBM, H, B);
    xattn_tf32_kernel<<<grid, kThreads, smem, st>>>(mq, mk, mv, prm);
    RD_LAUNCH_CHECK("xattn_tf32_kernel");
    return RD_OK;
}

}  // namespace rd
