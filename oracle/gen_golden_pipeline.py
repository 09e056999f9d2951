"""Fixtures that pin the SD2-architecture model code and the pipeline against the reference (build container).

Run through `python oracle/gen_golden.py --only pipeline`.  Weights are NOT stored: both sides fill their
parameters with tests/model_utils.deterministic_init_ (keyed by parameter name), which also proves the
state-dict keys are identical.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch


def run(da, rp, GOLD):
    sys.path.insert(0, os.path.dirname(GOLD.rstrip("/")))     # tests/
    import model_utils as mu
    from diffusers import AutoencoderKL, DDIMScheduler, UNet2DConditionModel

    # 1. state-dict keys/shapes of the full SD2 UNet and SD VAE (meta device: no memory)
    with torch.device("meta"):
        unet_full = UNet2DConditionModel(**mu.SD2_UNET)
        vae_full = AutoencoderKL(**mu.SD_VAE)
    keys = dict(unet={k: list(v.shape) for k, v in unet_full.state_dict().items()},
                vae={k: list(v.shape) for k, v in vae_full.state_dict().items()})
    with open(os.path.join(GOLD, "sd2_state_keys.json"), "w") as f:
        json.dump(keys, f)
    print("state keys:", len(keys["unet"]), len(keys["vae"]))

    # 2. tiny models, forward numerics
    torch.manual_seed(0)
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
    vae = mu.deterministic_init_(AutoencoderKL(**mu.TINY_VAE).eval(), seed=2)
    g = torch.Generator().manual_seed(5)
    out = {}
    for tag, (n, h, w) in {"even": (3, 8, 12), "odd": (3, 7, 13), "n4": (4, 6, 6)}.items():
        x = torch.randn(n, 8, h, w, generator=g)
        t = torch.full((n,), 999, dtype=torch.int64)
        ctx = torch.randn(1, 2, 32, generator=g)
        with torch.no_grad():
            y = unet(x, t, encoder_hidden_states=ctx, num_view=n).sample
            y_pf = unet(x, t, encoder_hidden_states=ctx.expand(n, -1, -1), num_view=None).sample
        out[f"unet_{tag}_x"], out[f"unet_{tag}_ctx"], out[f"unet_{tag}_y"] = x.numpy(), ctx.numpy(), y.numpy()
        out[f"unet_{tag}_y_perframe"] = y_pf.numpy()
    img = torch.rand(2, 3, 24, 40, generator=g) * 2 - 1
    with torch.no_grad():
        mom = vae.quant_conv(vae.encoder(img))
        z = mom[:, :4]
        dec = vae.decoder(vae.post_quant_conv(z))
    out["vae_img"], out["vae_moments"], out["vae_dec"] = img.numpy(), mom.numpy(), dec.numpy()

    # 3. whole pipeline on the tiny models (CPU fp32), with and without refinement
    sched = DDIMScheduler(**mu.DDIM_CFG)
    pipe = rp.RollingDepthPipeline(unet=unet, vae=vae, scheduler=sched, text_encoder=None, tokenizer=None)
    pipe.empty_text_embed = torch.randn(1, 2, 32, generator=g)
    out["pipe_text"] = pipe.empty_text_embed.numpy()
    frames = torch.rand(1, 10, 3, 40, 48, generator=g) * 2 - 1
    frames = torch.nn.functional.avg_pool2d(frames[0], 5, 1, 2)[None]        # smooth a little
    out["pipe_frames"] = frames.numpy()
    for tag, kw in {"norefine": dict(dilations=[1, 3], refine_step=0),
                    "refine": dict(dilations=[1, 2, 3], refine_step=2)}.items():
        with contextlib.redirect_stdout(io.StringIO()):
            res = pipe.forward(input_frames=frames.clone(), cap_dilation=True, snippet_lengths=[3], init_infer_steps=[1],
                               strides=[1], coalign_kwargs=dict(num_iterations=30), refine_snippet_len=3,
                               refine_start_dilation=3, generator=torch.Generator().manual_seed(0), verbose=False,
                               max_vae_bs=4, unload_snippet=False, **kw)
        out[f"pipe_{tag}_depth_pred"] = res.depth_pred.numpy()
        out[f"pipe_{tag}_coaligned"] = res.depth_coaligned.numpy()
        out[f"pipe_{tag}_rgb"] = res.input_rgb.numpy()
        for i, s in enumerate(res.snippet_ls):
            out[f"pipe_{tag}_snip{i}"] = s.numpy()
        print("pipeline", tag, tuple(res.depth_pred.shape), float(res.depth_pred.min()), float(res.depth_pred.max()))
    np.savez_compressed(os.path.join(GOLD, "models_pipeline.npz"), **out)
