"""Oracle for the whole snippet-inference path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

A device-agnostic restatement of RollingDepthPipeline.forward (rollingdepth/rollingdepth_pipeline.py:193-354,
356-463, 517-633, 635-740): one snippet per UNet call, one VAE chunk at a time, the oracle DDIM algebra and the
oracle aligner -- i.e. the reference's control flow, op for op.  `unet` / `vae` are passed in (any modules with the
diffusers call signatures); on CPU it doubles as the timed CPU baseline of bench.py.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import coalign as oc
from . import schedule as osch

SCALE = 0.18215


def encode_rgb(vae, rgb, max_bs):
    B, N = rgb.shape[:2]
    flat = rgb.reshape(B * N, *rgb.shape[-3:])
    lat = []
    for i in range(0, B * N, max_bs):
        mom = vae.quant_conv(vae.encoder(flat[i:i + max_bs]))
        lat.append(torch.chunk(mom, 2, dim=1)[0])
    return (torch.cat(lat) * SCALE).unflatten(0, (B, N))


def decode_depth(vae, lat, max_bs):
    B, N = lat.shape[:2]
    flat = lat.flatten(0, 1) / SCALE
    out = []
    for i in range(0, B * N, max_bs):
        out.append(vae.decoder(vae.post_quant_conv(flat[i:i + max_bs])))
    return torch.cat(out).mean(dim=1, keepdim=True).unflatten(0, (B, N))


def single_step(unet, rgb, depth, t, text):
    n = rgb.shape[1]
    x = torch.cat([rgb.flatten(0, 1), depth.flatten(0, 1)], dim=1)
    tt = torch.full((n,), int(t), dtype=torch.int64, device=x.device)
    return unet(x, tt, encoder_hidden_states=text, num_view=n).sample.unflatten(0, (1, n))


@torch.no_grad()
def forward(unet, vae, sched_cfg: Dict, text_embed, input_frames, dilations, cap_dilation=True, snippet_len=3,
            init_infer_step=1, coalign_kwargs: Optional[Dict] = None, refine_step=0, refine_snippet_len=3,
            refine_start_dilation=6, generator=None, max_vae_bs=4, timers: Optional[Dict] = None,
            aligner_factory=None):
    """Returns SimpleNamespace(depth_pred, depth_coaligned, snippet_ls, scales, translations, loss_history)."""
    import time
    tm = timers if timers is not None else {}

    def tick(name, t0):
        if input_frames.is_cuda:
            torch.cuda.synchronize()
        tm[name] = tm.get(name, 0.0) + time.time() - t0

    dev, dt = next(unet.parameters()).device, next(unet.parameters()).dtype
    dilations = list(dilations)
    N = input_frames.shape[1]
    if cap_dilation:
        dilations = [osch.cap_max_dilation(N, snippet_len, d) for d in dilations]
        refine_start_dilation = osch.cap_max_dilation(N, refine_snippet_len, refine_start_dilation)
    frames = input_frames.to(dt).to(dev)
    t0 = time.time(); rgb_latent = encode_rgb(vae, frames, max_vae_bs); tick("encode", t0)
    _, _, _, h, w = rgb_latent.shape
    noise = torch.randn((1, 4, h, w), device=dev, dtype=dt, generator=generator)
    noise = noise[:, None].expand(1, N, 4, h, w)
    text = text_embed.to(dev).to(dt)
    sch = osch.DDIMRef(**sched_cfg)

    snippet_ls = []
    for d in dilations:
        ts = sch.set_timesteps(init_infer_step)
        idx_ls = osch.get_snippet_indice(0, len(ts), N, snippet_len, d, d, 1)
        lat = []
        for idx in idx_ls:
            rgb_s, dep_s = rgb_latent[:, idx].clone(), noise[:, idx].clone()
            for t in ts:
                t0 = time.time(); pred = single_step(unet, rgb_s, dep_s, t, text); tick("unet", t0)
                dep_s = sch.step(pred, int(t), dep_s)
            lat.append(dep_s)
        t0 = time.time(); snippet_ls.append(decode_depth(vae, torch.cat(lat, 0), max_vae_bs)); tick("decode", t0)

    t0 = time.time()
    kw = {} if coalign_kwargs is None else dict(coalign_kwargs)
    aligner = (aligner_factory or oc.DepthAlignerRef)(device=dev, **kw)
    coal, scales, trans, hist = aligner.run([s for s in snippet_ls], dilations)
    coal = oc.renorm(coal.to(dev))
    tick("coalign", t0)

    if refine_step > 0:
        dl = encode_rgb(vae, coal[None].expand(1, N, 3, *coal.shape[-2:]), max_vae_bs)
        total = int(refine_step / 0.5)
        ts = sch.set_timesteps(total)
        ts = ts[int(len(ts) * 0.5):]
        new = sch.add_noise(dl, noise.clone(), int(ts[0]))
        for i_step, t in enumerate(ts):
            idx_ls = osch.get_snippet_indice(i_step, len(ts), N, refine_snippet_len, refine_start_dilation, 1, 1)
            old = new.clone()
            new = new * 0.0
            cnt = torch.zeros((1, N, 1, 1, 1), device=dev).int()
            for idx in idx_ls:
                t0 = time.time(); pred = single_step(unet, rgb_latent[:, idx], old[:, idx], t, text); tick("unet", t0)
                new[:, idx] += sch.step(pred, int(t), old[:, idx])
                cnt[:, idx] += 1
            new = new / cnt
        t0 = time.time(); depth_pred = decode_depth(vae, new, max_vae_bs); tick("decode", t0)
    else:
        depth_pred = coal
    return SimpleNamespace(depth_pred=depth_pred.squeeze(0) if depth_pred.dim() == 5 else depth_pred,
                           depth_coaligned=coal, snippet_ls=snippet_ls, scales=scales, translations=trans,
                           loss_history=hist, dilations=dilations)
