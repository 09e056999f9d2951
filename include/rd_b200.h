/*
 * rd_b200.h -- C ABI of librd_b200.so: hand-written sm_100a kernels for the
 * RollingDepth snippet-inference hot path.
 *
 * The reference (yizuo417/RollingDepth) is pure Python and has NO FFI for this
 * path: every kernel it runs comes from torch / xformers.  Each entry point
 * below therefore replaces a *call site* of the reference, cited per function;
 * the ctypes binding a maintainer would add on the reference side is shown in
 * INTEGRATION.md and lives in rollingdepth_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a
 *     parameter is explicitly named host_*;
 *   - strides are in ELEMENTS; the innermost (head_dim / pixel) axis is
 *     contiguous;
 *   - no allocation, no host synchronisation, stream-ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return 0 on success, a negative rd_status otherwise; rd_last_error()
 *     returns a thread-local human-readable message for the last failure;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns RD_ERR_CUDA.
 */
#ifndef RD_B200_H
#define RD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RD_B200_ABI_VERSION 1

typedef enum rd_dtype { RD_F32 = 0, RD_F16 = 1, RD_BF16 = 2 } rd_dtype;

typedef enum rd_status {
    RD_OK = 0,
    RD_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, misaligned pointer/stride */
    RD_ERR_UNSUPPORTED = -2,  /* shape/dtype the kernels are not instantiated for */
    RD_ERR_CUDA = -3,         /* CUDA runtime / driver error (message in rd_last_error) */
    RD_ERR_WORKSPACE = -4     /* caller-provided workspace too small */
} rd_status;

int rd_version(void);
const char* rd_last_error(void);
/* Compute capability of the current device as major*10+minor (100 on B200); <0 on error. */
int rd_device_arch(void);

/* ------------------------------------------------------------------ (1) cross-frame attention
 * Replaces F.scaled_dot_product_attention(q,k,v, attn_mask=None, dropout_p=0, is_causal=False)
 * at diffusers/src/diffusers/models/attention_processor.py:2251-2253 (AttnProcessor2_0) and
 * xformers.ops.memory_efficient_attention at :2036-2038 (XFormersAttnProcessor).
 *
 * o[b,s,h,:] = softmax_j( scale * q[b,s,h,:] . k[b,j,h,:] ) v[b,j,h,:]
 * q/o: [B, Sq, H, D] addressed as base + b*sb + s*ss + h*sh + d; k/v likewise with Sk.
 * This is the strided view the reference's to_q/to_k/to_v Linear layers already produce
 * ([B, S, H*D] viewed as [B, S, H, D]) -- no head-major copy is needed.
 * D must be 64 (every SD2 UNet level).  RD_F16/RD_BF16 run on tcgen05 tensor cores with
 * fp32 accumulation; RD_F32 (the `paper` preset / cfg1 tolerance path, |err| <= 1e-4) runs the error-compensated 3xTF32
 * tcgen05 kernel (hi/lo operand split, fp32 accumulation in TMEM); RD_XATTN_F32(1) selects the scalar fp32 FFMA kernel.
 * `flags`: 0 = default kernel; see RD_XATTN_* below (debug / A-B switches).
 */
/* share of the softmax exponentials evaluated by a degree-3 polynomial on the FMA pipes instead of MUFU.EX2
 * (the MUFU pipe is the first ceiling at head_dim 64): every n-th score pair; n = 1 disables, 0 = tuned default */
#define RD_XATTN_POLY(n) (((n) & 0xF) << 8)
/* query tiles per CTA: 0 = chosen by grid size (three 128-row tiles on many-wave grids of long sequences, else two),
 * 2 / 3 = force (A/B measurements; the parity tests run both) */
#define RD_XATTN_TILES(n) (((n) & 3) << 4)
/* fp32 inputs: 0 = default path, 1 = scalar FFMA kernel, 2 = error-compensated 3xTF32 tcgen05 kernel (A/B, parity tests of both) */
#define RD_XATTN_F32(n) (((n) & 3) << 14)
int rd_xattn_fwd(const void* q, const void* k, const void* v, void* o,
                 int B, int H, int Sq, int Sk, int D,
                 int64_t q_sb, int64_t q_ss, int64_t q_sh,
                 int64_t k_sb, int64_t k_ss, int64_t k_sh,
                 int64_t v_sb, int64_t v_ss, int64_t v_sh,
                 int64_t o_sb, int64_t o_ss, int64_t o_sh,
                 rd_dtype dtype, float scale, int flags, void* stream);

/* ------------------------------------------------------------------ (2) DDIM step
 * Replaces DDIMScheduler.step (diffusers/src/diffusers/schedulers/scheduling_ddim.py:342-468,
 * eta = 0, no thresholding) and add_noise (:471-495) as one element-wise launch each, with the
 * schedule scalars passed from the host (timesteps are known in advance, so the reference's
 * per-step host sync at :405-406 disappears).  Intermediate results are rounded to `dtype`
 * after every reference op so fp16 runs match the reference's op-by-op arithmetic.
 *   The four schedule scalars are fp32 values computed exactly like the reference does
 *   (alphas_cumprod[t] ** 0.5 etc. on 0-dim fp32 tensors).
 *   prediction_type: 0 = epsilon, 1 = sample, 2 = v_prediction.
 *   clip: <= 0 disables clip_sample, else clamp pred_x0 to [-clip, clip].
 */
int rd_ddim_step(const void* model_out, const void* sample, void* prev_sample, int64_t n,
                 float sqrt_alpha_t, float sqrt_beta_t, float sqrt_alpha_prev, float sqrt_one_minus_alpha_prev,
                 int prediction_type, float clip, rd_dtype dtype, void* stream);
int rd_ddim_add_noise(const void* original, const void* noise, void* out, int64_t n,
                      float sqrt_alpha_prod, float sqrt_one_minus_alpha_prod,
                      rd_dtype dtype, void* stream);

/* ------------------------------------------------------------------ (3) depth co-alignment
 * Replaces DepthAligner.run / optimize / merge_scaled_triplets
 * (rollingdepth/depth_aligner.py:68-120, 123-229, 231-262) and the pipeline's global
 * re-normalisation (rollingdepth/rollingdepth_pipeline.py:316-318).
 *
 * A problem has G groups (one per dilation d[g]), a uniform window w, N frames and
 * n[g] = N - (w-1)*d[g] snippets per group.  Snippet k of group g covers frames k + j*d[g].
 * Full-resolution snippets of group g are one dense array [n[g], w, H, W] (`snip[g]`).
 */
#define RD_MAX_GROUPS 8
#define RD_MAX_WINDOW 8

typedef struct rd_coalign_problem {
    int32_t G, w, N;
    int32_t H, W;              /* full resolution */
    int32_t border, factor;    /* crop and subsample (depth_aligner.py:83-92) */
    int32_t Hs, Ws;            /* subsampled size: ceil((H-2*border)/factor) etc. */
    int32_t P, P_pad;          /* Hs*Ws and its padding to a multiple of 8 */
    int32_t d[RD_MAX_GROUPS];
    int32_t n[RD_MAX_GROUPS];
    int32_t row0[RD_MAX_GROUPS]; /* first snippet row of group g in the packed arrays */
    int32_t n_total;           /* sum n[g] */
    int32_t dtype;             /* rd_dtype of the snippets */
} rd_coalign_problem;

/* Fill the derived fields (n, row0, n_total, Hs, Ws, P, P_pad) from G,w,N,H,W,border,factor,d. */
int rd_coalign_problem_init(rd_coalign_problem* p);

/* Global minimum over all snippets (depth_aligner.py:78): out_min[0] (fp32, exact: every
 * fp16/bf16 value is representable).  `scratch` >= rd_reduce_scratch_bytes(). */
int64_t rd_reduce_scratch_bytes(void);
int rd_global_min(const void* const* host_snip_ptrs, const rd_coalign_problem* p,
                  float* out_min, void* scratch, void* stream);

/* x_small[row, j, 0:P] = (snip - min)(in snippet dtype)[border::factor, border::factor]
 * (depth_aligner.py:79-92, :140-143); rows padded to P_pad with zeros.  x_small has the
 * snippet dtype and n_total*w*P_pad elements. */
int rd_coalign_prep(const void* const* host_snip_ptrs, const rd_coalign_problem* p,
                    const float* dev_min, void* x_small, void* stream);

typedef struct rd_coalign_hparams {
    float lr, beta1, beta2, eps;         /* Adam (depth_aligner.py:159): 1e-3, 0.5, 0.9, 1e-8 */
    float lmda2, lmda3;                  /* 1e-1, 1e1 (:205-209) */
    float depth_loss_weight, loss_scale; /* 1, 1 (:203, :211) */
} rd_coalign_hparams;

/* Kernel organisation of the optimisation pass (A/B measurements and the parity tests, which run all of them):
 * -1 = chosen per problem (default), 0 / 1 = pixel-packed kernel with 192 x 2 / 128 x 3 CTAs per SM,
 * 2 / 3 / 4 = contribution-packed kernel with 128 x 4 / 256 x 2 / 512 x 1 CTAs per SM, 5 / 6 = frame-persistent
 * contribution-packed kernel (one CTA per frame, two-stage ring) with 128 x 4 / 256 x 2.  Call before rd_coalign_workspace_bytes:
 * the workspace layout depends on it.  Returns the previous setting + 2 (>= 1), or a negative rd_status. */
int rd_coalign_set_config(int cfg);

/* Workspace size (bytes) for the optimisation entry points below. */
int64_t rd_coalign_workspace_bytes(const rd_coalign_problem* p, int n_iter);

/* Run `n_iter` iterations of the reference's optimisation loop (depth_aligner.py:217-227)
 * starting from state (s, t, m_s, m_t, v_s, v_t, step0); all state arrays are fp32 of
 * length n_total (group-major) and are updated in place.  history[3*it + {0,1,2}] =
 * (loss, target.min, target.max) evaluated BEFORE that iteration's update (:213).
 * One fused kernel per iteration, organised by frame: every subsampled prediction is
 * read exactly once per iteration (closed-form gradient, no autograd, Adam fused in). */
int rd_coalign_optimize(const void* x_small, const rd_coalign_problem* p, const rd_coalign_hparams* hp,
                        float* s, float* t, float* m_s, float* m_t, float* v_s, float* v_t,
                        int step0, int n_iter, float* history,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* The same loop split open for frame-sharded multi-GPU runs: every
 * rank holds all subsampled snippets and owns the frame range [f_begin, f_end).
 *   rd_coalign_begin : load state into the workspace, zero the exchange buffers.
 *   rd_coalign_pass  : pass `it` (0-based) over the owned frames; its prologue applies Adam
 *                      update number step0+it (skipped for it == 0) from the exchanged
 *                      gradient terms of pass it-1.
 *   rd_coalign_exchange : the per-(frame, contribution) gradient terms written by pass `it`
 *                      (*send, zeros outside the owned frames) and the buffer the next pass
 *                      reads (*recv).  Multi-GPU: ncclAllReduce(send -> recv, sum) after every
 *                      pass.  Single GPU: call rd_coalign_begin with single_rank = 1 and
 *                      send == recv (no copy).
 *   rd_coalign_end   : final Adam update for the snippets whose FIRST frame is owned (others
 *                      are left 0 so a sum-all-reduce assembles the result) and the per-rank
 *                      history partials: history[3*it+0] = sum of owned loss terms,
 *                      [1] = min, [2] = max over owned frames (reduce with sum / min / max).
 */
int rd_coalign_begin(const rd_coalign_problem* p, const float* s, const float* t, const float* m_s,
                     const float* m_t, const float* v_s, const float* v_t, int n_iter, int single_rank,
                     void* workspace, int64_t workspace_bytes, void* stream);
int rd_coalign_pass(const void* x_small, const rd_coalign_problem* p, const rd_coalign_hparams* hp,
                    int it, int step0, int f_begin, int f_end, void* workspace, void* stream);
int rd_coalign_exchange(const rd_coalign_problem* p, void* workspace, int it,
                        float** send, float** recv, int64_t* n_floats);
int rd_coalign_end(const rd_coalign_problem* p, const rd_coalign_hparams* hp, int n_iter, int step0,
                   int f_begin, int f_end, float* s, float* t, float* m_s, float* m_t, float* v_s,
                   float* v_t, float* history, void* workspace, void* stream);

/* merged[f, :, :] = mean over snippets covering f of ((snip - min) * s + t), all arithmetic in
 * the snippet dtype exactly as depth_aligner.py:246-262; also reduces the min / max of
 * `merged` into out_minmax[0..1] (fp32) for the pipeline renormalisation.
 * `scratch` >= rd_reduce_scratch_bytes(). */
int rd_merge(const void* const* host_snip_ptrs, const rd_coalign_problem* p, const float* dev_min,
             const float* s, const float* t, void* merged, float* out_minmax,
             void* scratch, void* stream);

/* In place: d = ((d - mn) / (mx - mn)) * 2 - 1 with the reference's op-by-op rounding
 * (rollingdepth_pipeline.py:316-318); minmax = {mn, mx} from rd_merge. */
int rd_renorm(void* depth, int64_t n, const float* dev_minmax, rd_dtype dtype, void* stream);

/* ------------------------------------------------------------------ (4) fused GroupNorm (+ SiLU), channels-last
 * Replaces the GroupNorm -> SiLU pairs in front of every convolution of the UNet / VAE resnets
 * (diffusers/src/diffusers/models/resnet.py:320-376: norm1/nonlinearity, norm2/nonlinearity), the
 * conv_norm_out/conv_act pairs (autoencoders/vae.py:160-165, 325-333; unets/unet_2d_condition.py:1312-1315)
 * and the act-free norms of Transformer2DModel (:418) and the VAE mid-block attention.
 * x, y: [N, HW, C] with C contiguous (= NCHW tensors in torch.channels_last memory format); gamma/beta: [C]
 * in the activation dtype (NULL = 1 / 0).  y = act(round(groupnorm(x))); statistics in fp32/fp64.
 * act: 0 = identity, 1 = SiLU.  workspace >= rd_groupnorm_workspace_bytes(N, G).
 * pre_add: optional [N, C] vector (activation dtype) added to x first: y = act(GN(round(x + pre_add[n, c]))) --
 * the bias of the preceding (bias-free) convolution plus the projected time embedding (resnet.py:341-352). */
int64_t rd_groupnorm_workspace_bytes(int N, int G);
int rd_groupnorm_nhwc(const void* x, void* y, const void* gamma, const void* beta, const void* pre_add,
                      int N, int HW, int C, int G, float eps, int act, rd_dtype dtype,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* y = x + bias[c] + vec[n, c] + residual in one vectorised pass over [N, HW, C] (any of bias / vec / residual may be
 * NULL; y may alias x): conv bias, time-embedding add and skip connection (resnet.py:341-376) without separate
 * element-wise launches.  Arithmetic in fp32, one rounding. */
int rd_bias_residual_nhwc(const void* x, const void* bias, const void* vec, const void* residual, void* y,
                          int N, int HW, int C, rd_dtype dtype, void* stream);

/* Nearest-neighbour resize of an NHWC activation, index rule of torch F.interpolate(mode="nearest")
 * (upsampling.py Upsample2D.forward). */
int rd_upsample_nearest_nhwc(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C,
                             rd_dtype dtype, void* stream);

/* Sub-pixel form of Upsample2D (upsampling.py Upsample2D.forward: nearest 2x upsampling, then the 3x3 convolution): the
 * convolution runs on the low-resolution tensor as one 2x2 convolution (padding 1) with 4 C output channels -- channel block
 * 2a + b holds output phase (a, b), its four tap weights are the sums of the 3x3 taps that read the same source pixel -- and this
 * entry point interleaves the phases and adds the bias:
 *     y[n, 2i + a, 2j + b, c] = t[n, i + a, j + b, (2a + b) C + c] + bias[c]
 * for t = [N, H + 1, W + 1, 4 C] and y = [N, 2 H, 2 W, C], both NHWC; bias may be NULL. */
int rd_phase_shuffle_bias_nhwc(const void* t, const void* bias, void* y, int N, int H, int W, int C,
                               rd_dtype dtype, void* stream);

/* GEGLU gate of the transformer feed-forward (diffusers/src/diffusers/models/activations.py GEGLU.forward, called from
 * attention.py FeedForward): y[m, d] = proj[m, d] * gelu(proj[m, D + d]) for a contiguous [M, 2 D] projection; exact
 * (erf) GELU, rounding as torch (gelu in fp32 -> dtype, product in fp32 -> dtype).  One pass instead of gelu + mul. */
int rd_geglu(const void* proj, void* y, int64_t M, int D, rd_dtype dtype, void* stream);

/* LayerNorm over the last axis of a contiguous [M, C] activation: the norm1 / norm2 / norm3 of every BasicTransformerBlock in front
 * of the cross-frame attention, the text cross-attention and the feed-forward (diffusers/src/diffusers/models/attention.py:
 * 447-448, 497-499, 520-521; torch.nn.functional.layer_norm).  Statistics in fp32 (two-pass), y = dtype(((x - mean) * rstd) *
 * gamma + beta), one read and one write.  C % (16 / sizeof(dtype)) == 0 and C <= 2048 (fp16/bf16) / 1024 (fp32), else
 * RD_ERR_UNSUPPORTED -- the caller keeps its library path. */
int rd_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t M, int C, float eps, rd_dtype dtype,
                 void* stream);

/* Row softmax of the VAE mid-block attention (one head of dim 512; diffusers/src/diffusers/models/unets/unet_2d_blocks.py
 * UNetMidBlock2D -> Attention -> F.scaled_dot_product_attention at attention_processor.py:2251): the fp32 logits come from a
 * cuBLAS batched GEMM; out[r, :] = dtype(softmax(scale * logits[r, :])) in ONE pass (scale, softmax and the cast to the
 * activation dtype were three element-wise passes over the [B, S, S] logits).  logits: contiguous [rows, cols] fp32,
 * cols % 4 == 0 and cols <= 10240 (RD_ERR_UNSUPPORTED otherwise -- the caller keeps its library path); out: [rows, cols]. */
int rd_softmax_rows(const float* logits, void* out, int64_t rows, int cols, float scale, rd_dtype out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RD_B200_H */
