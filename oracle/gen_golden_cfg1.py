"""Fixture for BASELINE.json configs[0] ("cfg1"): the REFERENCE pipeline itself (build container only), CPU fp32,
synthetic 16-frame 256x256 video, snippet length 3, dilations [1, 4], 1 denoise step, no refinement, 2000 co-alignment
iterations, random-init full SD2-architecture UNet (866 M) + SD VAE (84 M).

    python oracle/gen_golden.py --only cfg1          (about 3-4 minutes on 8 cores)

Weights are not stored (name-keyed generator, tests/model_utils.deterministic_init_); the frames come from
tests/model_utils.synth_video(16, 256, 256) and a float64 checksum of them is stored.  Stored: depth_pred
(== depth_coaligned without refinement) at full resolution, five decoded snippets at every second pixel, the aligner's
scales / translations / loss history, and the reference's own sensitivity floor (the same aligner run on the same
snippets perturbed by 1e-7: SURVEY.md section 7 "hard part 2").
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

GAIN = 0.7
SNIP_PICK = {0: (0, 7, 13), 1: (0, 7)}      # (group -> snippet rows) stored at [::2, ::2]


def absrel(a, b):
    d0, d1 = (a.float() + 1) / 2, (b.float() + 1) / 2
    return float(((d0 - d1).abs() / d1.clamp_min(1e-3)).mean())


def _reference_forward(rp, GOLD):
    """The reference pipeline on the cfg1 clip; returns (pipe, frames, result, seconds of init, seconds of forward)."""
    sys.path.insert(0, os.path.dirname(GOLD.rstrip("/")))     # tests/
    import model_utils as mu
    from diffusers import AutoencoderKL, DDIMScheduler, UNet2DConditionModel

    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.SD2_UNET).eval(), seed=11, gain=GAIN)
    vae = mu.deterministic_init_(AutoencoderKL(**mu.SD_VAE).eval(), seed=12, gain=GAIN)
    t_init = time.time() - t0
    pipe = rp.RollingDepthPipeline(unet=unet, vae=vae, scheduler=DDIMScheduler(**mu.DDIM_CFG), text_encoder=None,
                                   tokenizer=None)
    g = torch.Generator().manual_seed(21)
    pipe.empty_text_embed = torch.randn(1, 2, 1024, generator=g)
    frames = mu.synth_video(16, 256, 256, seed=0)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        res = pipe.forward(input_frames=frames.clone(), dilations=[1, 4], cap_dilation=True, snippet_lengths=[3],
                           init_infer_steps=[1], strides=[1], coalign_kwargs=None, refine_step=0, refine_snippet_len=3,
                           refine_start_dilation=6, generator=torch.Generator().manual_seed(0), verbose=False,
                           max_vae_bs=4, unload_snippet=False)
    t_fwd = time.time() - t0
    return pipe, frames, res, t_init, t_fwd


def _tail(hist, n=200):
    return sum(float(r[0]) for r in hist[-n:]) / n


def loss_floor(da, rp, GOLD, seeds=(5, 6, 7, 8)):
    """Adds the reference's own sensitivity floor of the LOSS CURVE to the existing fixture (other arrays untouched):
    the reference aligner on the reference's snippets against the same aligner on those snippets perturbed by 1e-7
    (one run per seed): relative difference of the mean loss over the last 200 iterations and of the first loss.
    tests/test_gpu_pipeline.py::test_cfg1_fp32_matches_reference_run derives its loss-level tolerance from these."""
    path = os.path.join(GOLD, "cfg1_reference.npz")
    old = dict(np.load(path))
    pipe, frames, res, _, t_fwd = _reference_forward(rp, GOLD)
    same = bool(np.array_equal(res.depth_pred.numpy(), old["depth_pred"]))
    dmax = float(np.abs(res.depth_pred.numpy() - old["depth_pred"]).max())
    print(f"cfg1 floor: reference forward re-run in {t_fwd:.1f} s; depth_pred identical to the stored fixture: {same} "
          f"(max-abs diff {dmax:.2e})")
    with contextlib.redirect_stdout(io.StringIO()):
        _, _, _, hist = da.DepthAligner(device=torch.device("cpu")).run([s.clone() for s in res.snippet_ls], [1, 4])
    tails, heads, scale_rels = [], [], []
    for sd in seeds:
        gq = torch.Generator().manual_seed(sd)
        pert = [s + 1e-7 * torch.randn(s.shape, generator=gq) for s in res.snippet_ls]
        with contextlib.redirect_stdout(io.StringIO()):
            _, _, _, hist_p = da.DepthAligner(device=torch.device("cpu")).run(pert, [1, 4])
        tails.append(abs(_tail(hist_p) - _tail(hist)) / abs(_tail(hist)))
        heads.append(abs(float(hist_p[0][0]) - float(hist[0][0])) / abs(float(hist[0][0])))
    old["floor_loss_tail_rel"] = np.array(tails)
    old["floor_loss_head_rel"] = np.array(heads)
    old["floor_loss_seeds"] = np.array(seeds)
    old["floor_loss_rerun_identical"] = np.array(same)
    np.savez_compressed(path, **old)
    print("cfg1 floor: loss tail rel (per seed)", ["%.2e" % t for t in tails], "; first-iteration loss rel",
          ["%.2e" % h for h in heads])


def run(da, rp, GOLD):
    pipe, frames, res, t_init, t_fwd = _reference_forward(rp, GOLD)
    assert torch.equal(res.depth_pred, res.depth_coaligned)
    out = dict(text=pipe.empty_text_embed.numpy(), frames_checksum=np.array(frames.double().sum().item()),
               depth_pred=res.depth_pred.numpy(),
               seconds=np.array([t_init, t_fwd]), cores=np.array(os.cpu_count() or 1))
    for gi, rows in SNIP_PICK.items():
        out[f"snip{gi}_shape"] = np.array(res.snippet_ls[gi].shape)
        for r in rows:
            out[f"snip{gi}_{r}"] = res.snippet_ls[gi][r, :, :, ::2, ::2].numpy()
    # the aligner alone, on the reference's snippets: parameters, history and the reference's own sensitivity floor
    with contextlib.redirect_stdout(io.StringIO()):
        al = da.DepthAligner(device=torch.device("cpu"))
        merged, scales, trans, hist = al.run([s.clone() for s in res.snippet_ls], [1, 4])
        gq = torch.Generator().manual_seed(5)
        pert = [s + 1e-7 * torch.randn(s.shape, generator=gq) for s in res.snippet_ls]
        merged_p, scales_p, _, _ = da.DepthAligner(device=torch.device("cpu")).run(pert, [1, 4])

    def renorm(d):
        d = d - d.min()
        d = d / d.max()
        return d * 2 - 1
    assert torch.equal(renorm(merged.detach()), res.depth_coaligned)
    out["floor_absrel"] = np.array(absrel(renorm(merged_p.detach()), renorm(merged.detach())))
    out["floor_scale_rel"] = np.array(max(float(((a - b).abs() / b.abs()).max()) for a, b in zip(scales_p, scales)))
    for i in range(2):
        out[f"scale{i}"] = scales[i].detach().reshape(-1).numpy()
        out[f"trans{i}"] = trans[i].detach().reshape(-1).numpy()
    out["hist"] = np.array(hist, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "cfg1_reference.npz"), **out)
    print(f"cfg1: reference forward {t_fwd:.1f} s on {os.cpu_count()} cores (model init {t_init:.1f} s); "
          f"floor AbsRel {float(out['floor_absrel']):.2e}, floor scale rel {float(out['floor_scale_rel']):.2e}; "
          f"depth range [{float(res.depth_pred.min()):.3f}, {float(res.depth_pred.max()):.3f}]")
