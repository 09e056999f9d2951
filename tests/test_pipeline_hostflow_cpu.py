"""The product pipeline's HOST control flow on CPU against the REFERENCE pipeline's own outputs.

`RollingDepthPipeline.init_snippet_infer` (dilation groups, flattened snippet list, balanced batches of several snippets per
UNet call, decode straight into the gather buffer), `refine` (annealed windows, batched, `_average_windows`), `encode_rgb` /
`decode_depth` and `B200XFrameAttnProcessor` are host code around four calls into librd_b200.so.  Here those calls
(`ops.xattn_fwd`, `ops.ddim_step`, `ops.ddim_add_noise`; the aligner as a whole) are replaced by the oracle's CPU arithmetic --
test-only: the product has no CPU path and `forward()` itself refuses a CPU device -- and the stages are chained exactly as
`forward` chains them.  The results are compared with tests/golden/models_pipeline.npz, i.e. with what the reference's
`forward` (rollingdepth/rollingdepth_pipeline.py:194-353, one snippet per UNet call) produced for the same frames, weights,
text embedding and noise generator.  The kernels themselves are pinned by the -m gpu tier.
"""
import os

import numpy as np
import pytest
import torch

import model_utils as mu
from oracle import coalign as oc
from oracle import xframe_attention as oxa


@pytest.fixture()
def library_calls_stubbed_by_the_oracle(monkeypatch):
    from rollingdepth_b200 import _lib as L
    from rollingdepth_b200 import ops
    calls = dict(xattn=[], ddim_step=0, add_noise=0)

    def xattn(q, k, v, scale=None, out=None, flags=0):
        calls["xattn"].append(tuple(q.shape))
        o = oxa.sdpa_ref(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float(), scale=scale)
        return o.transpose(1, 2).to(q.dtype)

    def ddim_step(model_output, sample, sa, sb, sap, sbp, prediction_type, clip=0.0, out=None):
        # scheduling_ddim.py:405-452 (eta = 0) with the host-side scalars, v-prediction only (RollingDepth's scheduler)
        assert prediction_type == "v_prediction" and clip == 0.0
        sa, sb, sap, sbp = [torch.tensor(s, dtype=torch.float32) for s in (sa, sb, sap, sbp)]
        x0 = sa * sample - sb * model_output
        eps = sa * model_output + sb * sample
        y = sap * x0 + sbp * eps
        calls["ddim_step"] += 1
        if out is None:
            return y
        out.copy_(y)
        return out

    def add_noise(original, noise, sa, sb):
        calls["add_noise"] += 1
        return torch.tensor(sa, dtype=original.dtype) * original + torch.tensor(sb, dtype=original.dtype) * noise

    monkeypatch.setattr(ops, "xattn_fwd", xattn)
    monkeypatch.setattr(ops, "ddim_step", ddim_step)
    monkeypatch.setattr(ops, "ddim_add_noise", add_noise)
    monkeypatch.setattr(L, "require_cuda", lambda *a, **k: None)
    return calls


def _absrel(a, b):
    d0, d1 = (a.float() + 1) / 2, (b.float() + 1) / 2
    return float(((d0 - d1).abs() / d1.clamp_min(1e-3)).mean())


@pytest.mark.parametrize("tag,dil,refine", [("norefine", [1, 3], 0), ("refine", [1, 2, 3], 2)])
@pytest.mark.parametrize("snippet_batch,max_vae_bs", [(1, 4), (4, 4), (16, 36)])
def test_host_control_flow_reproduces_the_reference_run(golden_dir, library_calls_stubbed_by_the_oracle, tag, dil, refine,
                                                        snippet_batch, max_vae_bs):
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    from rollingdepth_b200.sd2.unet import UNet2DConditionModel
    from rollingdepth_b200.sd2.vae import AutoencoderKL
    calls = library_calls_stubbed_by_the_oracle
    g = np.load(os.path.join(golden_dir, "models_pipeline.npz"))
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=snippet_batch)     # installs the processor
    pipe.empty_text_embed = torch.from_numpy(g["pipe_text"])
    frames = torch.from_numpy(g["pipe_frames"])
    N = frames.shape[1]
    with torch.no_grad():
        # ---- the stages of forward(), in forward()'s order (rollingdepth_pipeline.py:244-345)
        dil = [pipe.cap_max_dilation(seq_len=N, snippet_len=3, dilation=d, verbose=False) for d in dil]
        start = pipe.cap_max_dilation(seq_len=N, snippet_len=3, dilation=3, verbose=False)
        rgb_latent = pipe.encode_rgb(frames, max_batch_size=max_vae_bs)
        _, _, _, h, w = rgb_latent.shape
        noise = torch.randn((1, 4, h, w), generator=torch.Generator().manual_seed(0))
        init_noise = noise[:, None].expand(1, N, 4, h, w)
        snips = pipe.init_snippet_infer(rgb_latent=rgb_latent, init_noise=init_noise, dilations=dil, cap_dilation=True,
                                        snippet_lengths=[3] * len(dil), init_infer_steps=[1] * len(dil), strides=[1] * len(dil),
                                        max_vae_bs=max_vae_bs, unload_snippet=False, verbose=False)
        assert len(snips) == len(dil)
        for i, s in enumerate(snips):
            ref = torch.from_numpy(g[f"pipe_{tag}_snip{i}"])
            assert s.shape == ref.shape and float((s - ref).abs().max()) < 2e-5, (i, float((s - ref).abs().max()))
        # cross-frame attention really ran over (snippets per call) x (3 frames x tokens): the batched call is the product's
        n_snip = sum(N - 2 * d for d in dil)
        n_calls = -(-n_snip // snippet_batch)                       # all dilations share (length, steps): ONE flattened list
        assert calls["ddim_step"] == n_calls                        # one UNet call + one DDIM step per balanced batch
        assert max(shape[0] for shape in calls["xattn"]) == -(-n_snip // n_calls)       # e.g. 20 snippets, batch 16 -> 10 + 10
        coal, _, _, _ = oc.DepthAlignerRef(num_iterations=30).run(snips, dil)
        coal = oc.renorm(coal)
        # the aligner (closed-form oracle, 30 sign-gradient Adam steps) amplifies 1e-6 snippet differences: same tolerance as the
        # oracle-pipeline test and the north star (1e-3); measured 1.3e-4
        assert _absrel(coal, torch.from_numpy(g[f"pipe_{tag}_coaligned"])) <= 1e-3
        if refine:
            dl = pipe.encode_rgb(coal[None].expand(1, N, 3, *coal.shape[-2:]), max_batch_size=max_vae_bs)
            new = pipe.refine(rgb_latent=rgb_latent, depth_latents=dl, init_noise=init_noise, refine_step=refine, snippet_len=3,
                              start_dilation=start, verbose=False)
            assert calls["add_noise"] == 1
            depth_pred = pipe.decode_depth(new, max_batch_size=max_vae_bs).squeeze(0)
        else:
            depth_pred = coal
    ref_pred = torch.from_numpy(g[f"pipe_{tag}_depth_pred"])
    assert depth_pred.shape == ref_pred.shape
    assert _absrel(depth_pred, ref_pred) <= 1e-3          # north-star tolerance
