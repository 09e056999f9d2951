// Cross-frame self-attention forward for sm_100a: flash-style, head_dim 64, fp16/bf16 in,
// fp32 accumulate.  Replaces the fused-attention library call of the reference's processors
// (diffusers/src/diffusers/models/attention_processor.py:2251-2253 and :2036-2038).
//
// At head_dim 64 one 128x128 score tile costs 512 tensor-pipe cycles (QK^T + PV) but 1024 MUFU cycles
// (16 exponentials / clk / SM, tools/probes/pipe_rate.cu), so the kernel is organised around keeping the MUFU pipe
// busy: every softmax warp is software-pipelined so that its own TMEM loads, row maxima and barrier traffic sit in
// the shadow of its exponentials, and two such warps share each SM sub-partition.
//
// One CTA per SM = TWO 128-row query tiles of one (batch, head) that share every K/V tile.  The softmax works in
// STEPS of 64 keys (half a 128-key K/V tile) so that scores and probabilities can be double-buffered in TMEM:
//   per query tile t (256 columns at 256 t):  S[2] fp32 64 columns each, O fp32 64 columns, P[2] packed 16-bit 32 each.
// Warp roles (384 threads; setmaxnreg moves registers from the third warpgroup to the softmax warpgroups):
//   warps 0-3  softmax of query tile 0, warps 4-7 softmax of query tile 1.  Each thread owns one query row
//              (= one TMEM lane).  Step j: issue the TMEM load of S(j+1) (tcgen05.ld, asynchronous), evaluate the
//              exponentials of S(j) from registers, store P(j) (tcgen05.st), then complete the load of S(j+1), free
//              its TMEM buffer, take its row maximum, and only then wait for the P store and release PV(j).
//              O is rescaled lazily (only when the running maximum grows by more than 2^8), which is the only place
//              a softmax warp has to wait for a just-issued MMA.
//   warps 8,9  tcgen05.mma issuers, one elected thread per query tile (warp 8 also owns the TMEM allocation):
//              S_t(j) = Q_t K(j)^T (SS mode, both K-major, SWIZZLE_128B, N = 64), issued two steps ahead;
//              O_t += P_t(j) V(j) (A = P_t from TMEM, B = V MN-major from shared memory).  K/V stages are
//              released when both issuers have committed.
//   warp 10    TMA producer: both Q tiles once, then K/V tiles of 128 keys (two steps) through 3-stage rings.
// Element-wise softmax arithmetic uses the packed fp32x2 instructions of sm_100 (FFMA2 / FADD2: one issue slot for
// two results), and a tunable share of the exponentials can be evaluated by a degree-3 polynomial on the FMA pipe.
// Earlier organisations (one query tile per CTA x 2 CTAs/SM, 128-key steps with single-buffered S, forced ping-pong
// of the warpgroups, one in-order issuer) are archived with their measurements under tools/probes/attn_variants/.
#include <cuda.h>
#include <math.h>
#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tensormap.cuh"

namespace rd {

using namespace ptx;

#ifdef RD_XATTN_TRACE
// phase cycle counters of the first warp of each softmax warpgroup (summed over CTAs), read by tools/probes/xattn_bench.cu
__device__ unsigned long long rd_xattn_trace[16];
// timeline of one CTA (block 0,0,0): actors = softmax warp 0, softmax warp 4, issuer 0, issuer 1; steps 64..95
__device__ long long rd_xattn_tl[4][32][8];
#ifdef RD_XATTN_TIMELINE
#define RD_TL(actor, j, i) if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) >= 64 && (j) < 96) rd_xattn_tl[actor][(j) - 64][i] = clock64();
#else
#define RD_TL(actor, j, i)
#endif
#define RD_TRACE_DECL long long tr_t = clock64(); unsigned long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; int tr_j = 0;
#define RD_TRACE(i) { const long long tr_n = clock64(); tr_acc[i] += (unsigned long long)(tr_n - tr_t); tr_t = tr_n; if ((threadIdx.x & 127) == 0) { RD_TL(threadIdx.x >> 7, tr_j, i) } if (i == 5) ++tr_j; }
#define RD_TRACE_FLUSH if ((threadIdx.x & 127) == 0) { for (int tr_i = 0; tr_i < 7; ++tr_i) atomicAdd(&rd_xattn_trace[tr_i], tr_acc[tr_i]); atomicAdd(&rd_xattn_trace[8 + (threadIdx.x >> 7)], 1ull); }
#else
#define RD_TL(actor, j, i)
#define RD_TRACE_DECL
#define RD_TRACE(i)
#define RD_TRACE_FLUSH
#endif

constexpr int kBM = 128;       // rows per query tile
constexpr int kNQ = 2;         // query tiles per CTA
constexpr int kBN = 128;       // keys per K/V tile (one TMA box)
constexpr int kBK = 64;        // keys per softmax / MMA step
constexpr int kD = 64;         // head dim
constexpr int kStages = 3;
constexpr int kTileBytes = kBN * kD * 2;   // 16384
constexpr int kStepBytes = kBK * kD * 2;   // 8192: second step of a K/V tile
constexpr int kSoftmaxThreads = kNQ * 128;
constexpr int kThreads = kSoftmaxThreads + 128;   // third warpgroup: two MMA warps, TMA warp, one idle warp (register donors)
constexpr int kMmaWarp = kNQ * 4, kTmaWarp = kNQ * 4 + 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColTile = 256, kColS = 0, kColO = 128, kColP = 192;   // S: + 64 b, P: + 32 b
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
// barrier indices; S_FULL / S_FREE / P_FULL / O_READY are indexed [2 t + buffer]
constexpr int Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + kStages, V_FULL = K_EMPTY + kStages, V_EMPTY = V_FULL + kStages,
              S_FULL = V_EMPTY + kStages, S_FREE = S_FULL + 2 * kNQ, P_FULL = S_FREE + 2 * kNQ, O_READY = P_FULL + 2 * kNQ;
static_assert(O_READY + 2 * kNQ <= 32, "barrier table too small");

struct XattnParams {
    void* o;
    long long o_sb, o_ss, o_sh;
    int Sq, Sk, n_kv_tiles, n_steps;
    float scale_log2;   // scale * log2(e)
    int wg1_delay;      // cycles the second softmax warpgroup starts late (de-phases the two warps of a sub-partition)
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
// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2: one issue slot, two lanes of work)
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
// relative error 7.5e-5, below the fp16/bf16 rounding of P), exponent patched in with one shift-add.
// Costs ~10 issue slots per element (pipe_rate probe) against 8 MUFU cycles, so only a small share pays off.
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
    p0 = __int_as_float((__float_as_int(t0) << 23) + __float_as_int(q0));
    p1 = __int_as_float((__float_as_int(t1) << 23) + __float_as_int(q1));
}

// The destination registers of an asynchronous tcgen05.ld are only valid after tcgen05.wait::ld; this empty
// volatile asm (ordered after the wait) makes every later use of the values depend on it.
__device__ __forceinline__ void pin32(uint32_t (&r)[32]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                      "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

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
template <typename T, int POLY, bool MIDLD>
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
        for (int i = 0; i < 2 * kNQ; ++i) {
            mbar_init(&bars[S_FULL + i], 1);
            mbar_init(&bars[S_FREE + i], 4);       // one elected lane per softmax warp
            mbar_init(&bars[P_FULL + i], 4);
            mbar_init(&bars[O_READY + i], 1);
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
        } else if (warp < kMmaWarp + kNQ) {
        // ===================================================== MMA issuer of query tile t (uniform control flow)
        {
            const bool leader = elect_one();
            const uint32_t tmu = __shfl_sync(0xffffffffu, tmem_base, 0);
            constexpr uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
            constexpr uint32_t idesc_qk = make_idesc_f16(fmt, 0, 0, kBM, kBK);
            constexpr uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, kBM, kD);   // B = V is MN-major
            const int t = warp - kMmaWarp;
            const uint32_t tS = tmu + t * kColTile + kColS, tO = tmu + t * kColTile + kColO,
                           tP = tmu + t * kColTile + kColP;
            const uint64_t dq = make_smem_desc_sw128(smem_u32(smem + XattnSmem::q + t * kTileBytes), 16, 1024);

            auto issue_qk = [&](int i) {      // S(i) = Q K(i)^T into S buffer i & 1
                const int tile = i >> 1, st = tile % kStages, b = i & 1;
                mbar_wait(&bars[K_FULL + st], (tile / kStages) & 1);
                if (i >= 2) mbar_wait(&bars[S_FREE + 2 * t + b], ((i - 2) >> 1) & 1);   // softmax has S(i-2) in registers
                tc_fence_after();
                const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + XattnSmem::k + st * kTileBytes + b * kStepBytes), 16, 1024);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kD / 16; ++kk)             // +32 bytes along K inside the 128B swizzle row
                        mma_f16_ss(tS + b * kBK, dq + 2 * kk, dk + 2 * kk, idesc_qk, kk > 0);
                    if (b == 1 || i == n_steps - 1) mma_commit(&bars[K_EMPTY + st]);
                    mma_commit(&bars[S_FULL + 2 * t + b]);
                }
                __syncwarp();
            };

            // Issue order = the order in which the softmax warps produce the events: S(i) is in registers (S_FREE)
            // half a step before P(i-1) is stored (P_FULL), so QK runs three steps ahead of PV.  Because the tensor
            // pipe executes one thread's MMAs in order, S_FULL(j+1) also tells the softmax warps that PV(j-2) is done.
            mbar_wait(&bars[Q_FULL], 0);
            issue_qk(0);
            if (n_steps > 1) issue_qk(1);
            if (n_steps > 2) issue_qk(2);
            for (int j = 0; j < n_steps; ++j) {
                const int tile = j >> 1, st = tile % kStages, b = j & 1;
                RD_TL(2 + t, j, 0)
                mbar_wait(&bars[V_FULL + st], (tile / kStages) & 1);
                RD_TL(2 + t, j, 1)
                mbar_wait(&bars[P_FULL + 2 * t + b], (j >> 1) & 1);
                RD_TL(2 + t, j, 2)
                tc_fence_after();
                const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + XattnSmem::v + st * kTileBytes + b * kStepBytes), 16, 1024);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < kBK / 16; ++kk)            // V advances 16 key rows = 2048 bytes per K step
                        mma_f16_ts(tO, tP + b * (kBK / 2) + 8 * kk, dv + (uint64_t)(128 * kk), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                    if (b == 1 || j == n_steps - 1) mma_commit(&bars[V_EMPTY + st]);
                    mma_commit(&bars[O_READY + 2 * t + b]);
                }
                __syncwarp();
                RD_TL(2 + t, j, 3)
                if (j + 3 < n_steps) issue_qk(j + 3);
                RD_TL(2 + t, j, 4)
            }
        }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ===================================================== softmax warps (one query row per thread)
        const int t = warp >> 2;                       // query tile of this warpgroup
        const int row = (warp & 3) * 32 + lane;        // row inside the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_off + t * kColTile + kColS, tO = tmem_base + lane_off + t * kColTile + kColO,
                       tP = tmem_base + lane_off + t * kColTile + kColP;
        uint64_t* s_full = &bars[S_FULL + 2 * t];
        uint64_t* s_free = &bars[S_FREE + 2 * t];
        uint64_t* p_full = &bars[P_FULL + 2 * t];
        uint64_t* o_ready = &bars[O_READY + 2 * t];
        const float sl2 = prm.scale_log2;
        const uint64_t sl2_2 = pack2(sl2, sl2);
        float m_used;               // raw-score units
        float mt_next = 0.f;        // row maximum of the prefetched step
        uint64_t l2 = 0ull;         // two partial row sums (fp32 x 2)
        RD_TRACE_DECL

        // S(j) lives in (xa, xb), S(j+1) is prefetched into (ya, yb); the loop is unrolled by two so that the register
        // arrays keep static names.
        uint32_t s0a[32], s0b[32], s1a[32], s1b[32];

        // ---- prologue: S(0) -> registers, row maximum
        if (t == 1 && prm.wg1_delay > 0) {
            const long long t_start = clock64();
            while (clock64() - t_start < prm.wg1_delay) {}
        }
        mbar_wait(&s_full[0], 0);
        tc_fence_after();
        tmem_ld32(tS + 0, s0a);
        tmem_ld32(tS + 32, s0b);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[0]);
        if (prm.Sk < kBK) mask_tail64(s0a, s0b, prm.Sk);
        m_used = row_max64(s0a, s0b);

        auto step = [&](int j, uint32_t (&xa)[32], uint32_t (&xb)[32], uint32_t (&ya)[32], uint32_t (&yb)[32]) {
            const int b = j & 1, nb = b ^ 1;
            const bool has_next = j + 1 < n_steps;
            // -- prefetch S(j+1): asynchronous TMEM load, completed after the exponentials below
            if (has_next && !MIDLD) {
                mbar_wait(&s_full[nb], ((j + 1) >> 1) & 1);
                tc_fence_after();
                RD_TRACE(0)
                tmem_ld32(tS + nb * kBK + 0, ya);
                tmem_ld32(tS + nb * kBK + 32, yb);
            }
            // -- P buffer b is free: PV(j-2) precedes QK(j+1) in the issuer's in-order MMA stream, whose completion
            //    (S_FULL(j+1)) was just observed; the last step has no S_FULL to lean on
            if ((!has_next || MIDLD) && j >= 2) {
                mbar_wait(&o_ready[b], ((j - 2) >> 1) & 1);
                tc_fence_after();
            }
            RD_TRACE(1)
            // -- exponentials of S(j), P(j) -> TMEM
            const float mb = m_used * sl2;
            const uint64_t nmb2 = pack2(-mb, -mb);
            uint64_t ls2[2] = {l2, 0ull};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const uint32_t r0 = h == 0 ? xa[i] : xb[i], r1 = h == 0 ? xa[i + 1] : xb[i + 1];
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
                tmem_st16(tP + b * (kBK / 2) + h * 16, pk);
                if (MIDLD && h == 0 && has_next) {
                    mbar_wait(&s_full[nb], ((j + 1) >> 1) & 1);
                    tc_fence_after();
                    tmem_ld32(tS + nb * kBK + 0, ya);
                    tmem_ld32(tS + nb * kBK + 32, yb);
                }
            }
            l2 = fadd2(ls2[0], ls2[1]);
            RD_TRACE(2)
            // -- S(j+1) has landed: free its TMEM buffer (QK(j+3) may overwrite it), row maximum
            if (has_next) {
                tmem_ld_wait();
                pin32(ya);
                pin32(yb);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[nb]);
                const int valid = prm.Sk - (j + 1) * kBK;
                if (valid < kBK) mask_tail64(ya, yb, valid);
                mt_next = row_max64(ya, yb);
            }
            RD_TRACE(3)
            // -- P(j) is in TMEM: release PV(j)
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[b]);
            RD_TRACE(4)
            // -- lazy rescaling for step j+1 (rare after the first few steps): needs PV(j) complete
            if (has_next) {
                const float m_new = fmaxf(m_used, mt_next);
                const bool need = (m_new - m_used) * sl2 > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    mbar_wait(&o_ready[b], (j >> 1) & 1);
                    tc_fence_after();
                    const float f = ex2((m_used - m_new) * sl2);
                    float la, lb;
                    unpack2(l2, la, lb);
                    l2 = pack2(la * f, lb * f);
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
            }
            RD_TRACE(5)
        };

        for (int j = 0; j < n_steps; j += 2) {
            step(j, s0a, s0b, s1a, s1b);
            if (j + 1 < n_steps) step(j + 1, s1a, s1b, s0a, s0b);
        }
        RD_TRACE_FLUSH

        // ------------------------------------------------- epilogue: O / l -> global
        mbar_wait(&o_ready[(n_steps - 1) & 1], ((n_steps - 1) >> 1) & 1);   // the tensor pipe is in order: all PV done
        tc_fence_after();
        uint32_t orr[2][32];
        tmem_ld32(tO + 0, orr[0]);
        tmem_ld32(tO + 32, orr[1]);
        tmem_ld_wait();
        float la, lb;
        unpack2(l2, la, lb);
        const float inv_l = 1.0f / (la + lb);
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
int xattn_f32_launch(const float* q, const float* k, const float* v, float* o, int B, int H, int Sq, int Sk,
                     const int64_t* qs, const int64_t* ks, const int64_t* vs, const int64_t* os, float scale,
                     cudaStream_t st);

template <typename T, int POLY, bool MIDLD>
static int launch_k(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    cudaStream_t st) {
    auto kern = xattn_fwd_kernel<T, POLY, MIDLD>;
    const int smem = XattnSmem::bytes + 1024;
    static bool configured = false;      // per instantiation; cudaFuncSetAttribute is not a stream operation
    if (!configured) {
        RD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    kern<<<grid, kThreads, smem, st>>>(mq, mk, mv, prm);
    RD_LAUNCH_CHECK("xattn_fwd_kernel");
    return RD_OK;
}

constexpr int kDefaultPoly = 1;   // polynomial path off by default (attention log)

template <typename T>
static int launch_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const XattnParams& prm, dim3 grid,
                    int flags, cudaStream_t st) {
    int poly = (flags >> 8) & 0xF;       // RD_XATTN_POLY(n): 0 = default, 1 = off, n >= 2 = every n-th pair
    if (poly == 0) poly = kDefaultPoly;
    if (flags & 1) return launch_k<T, 0, true>(mq, mk, mv, prm, grid, st);
    switch (poly) {
        case 1: return launch_k<T, 0, false>(mq, mk, mv, prm, grid, st);
        case 4: return launch_k<T, 4, false>(mq, mk, mv, prm, grid, st);
        case 8: return launch_k<T, 8, false>(mq, mk, mv, prm, grid, st);
        default: set_error("RD_XATTN_POLY(%d) is not instantiated (1, 2, 4, 8, 15)", poly); return RD_ERR_UNSUPPORTED;
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
    prm.Sq = Sq; prm.Sk = Sk; prm.n_kv_tiles = (Sk + kBN - 1) / kBN; prm.n_steps = (Sk + kBK - 1) / kBK;
    prm.scale_log2 = scale * 1.4426950408889634f;
    prm.wg1_delay = ((flags >> 16) & 0xFFF) * 8;
    dim3 grid((Sq + kNQ * kBM - 1) / (kNQ * kBM), H, B);
    if (dtype == RD_F16) return launch_t<__half>(mq, mk, mv, prm, grid, flags, st);
    return launch_t<__nv_bfloat16>(mq, mk, mv, prm, grid, flags, st);
}
