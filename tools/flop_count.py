"""Convolution / linear FLOPs of one VAE-encoder, UNet and VAE-decoder call at a given frame size, counted with forward
hooks on meta tensors (no compute; runs on the CPU box).  Used for the step-level roofline paragraph of DESIGN.md.

    python tools/flop_count.py --height 432 --width 768 --frames 150 --dilations 1 25
"""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def count(module, *inputs, **kw):
    tot = {"conv": 0, "linear": 0}
    hooks = []

    def conv_hook(m, inp, out):
        kh, kw_ = m.kernel_size
        tot["conv"] += 2 * out.numel() * (m.in_channels // m.groups) * kh * kw_

    def lin_hook(m, inp, out):
        tot["linear"] += 2 * out.numel() * m.in_features

    def up_hook(m, inp, out):
        # exact-2x Upsample2D runs in sub-pixel form: 16 instead of 36 multiply-adds per source pixel and channel pair
        if tuple(out.shape[-2:]) == (2 * inp[0].shape[-2], 2 * inp[0].shape[-1]):
            saved = 2 * out.numel() * m.conv.in_channels * 9 * 20 // 36
            tot["conv"] -= saved
            tot["subpixel_saved"] += saved

    from rollingdepth_b200.sd2.layers import Upsample2D
    tot["subpixel_saved"] = 0
    for m in module.modules():
        if isinstance(m, Upsample2D):
            hooks.append(m.register_forward_hook(up_hook))
        if isinstance(m, nn.Conv2d):
            hooks.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.Linear):
            hooks.append(m.register_forward_hook(lin_hook))
    with torch.no_grad():
        module(*inputs, **kw)
    for h in hooks:
        h.remove()
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=432)
    ap.add_argument("--width", type=int, default=768)
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--dilations", type=int, nargs="+", default=[1, 25])
    ap.add_argument("--peak", type=float, default=1395.7, help="sustained dense bf16 TFLOP/s (MEASURED_PEAKS.json)")
    a = ap.parse_args()
    from rollingdepth_b200.sd2.vae import SD_VAE_CONFIG, AutoencoderKL
    with torch.device("meta"):
        vae = AutoencoderKL(**SD_VAE_CONFIG)
        x = torch.empty(1, 3, a.height, a.width)
        enc = count(vae.encoder, x)
        z = torch.empty(1, 4, a.height // 8, a.width // 8)
        dec = count(vae.decoder, z)
    from rollingdepth_b200.sd2.unet import SD2_UNET_CONFIG, UNet2DConditionModel
    h, w = a.height // 8, a.width // 8
    with torch.device("meta"):
        unet = UNet2DConditionModel(**SD2_UNET_CONFIG)
        un = count(unet, torch.empty(3, 8, h, w), torch.tensor(999), torch.empty(1, 2, 1024), num_view=3)
    # cross-frame self-attention (4 * Sq * Sk * D * heads per layer, Sq = Sk = 3 * tokens of the level): 5 layers on
    # each of the three attention levels (2 down + 3 up) and 1 in the mid block, head_dim 64, heads 5 / 10 / 20 / 20
    attn = 0
    for lvl, (heads, layers) in enumerate(((5, 5), (10, 5), (20, 5), (20, 1))):
        hh, ww = h, w
        for _ in range(lvl):
            hh, ww = (hh + 1) // 2, (ww + 1) // 2
        attn += layers * 4 * (3 * hh * ww) ** 2 * 64 * heads
    n_snip = sum(max(a.frames - 2 * d, 0) for d in a.dilations)
    print(f"UNet, one 3-frame snippet: {(un['conv'] + un['linear']) / 1e12:.3f} TFLOP conv + linear "
          f"({un['subpixel_saved'] / 1e12:.3f} saved by sub-pixel upsampling), {attn / 1e12:.3f} TFLOP "
          f"cross-frame self-attention; x {n_snip} snippets = {n_snip * (un['conv'] + un['linear']) / 1e15:.3f} + "
          f"{n_snip * attn / 1e15:.3f} PFLOP = {n_snip * (un['conv'] + un['linear']) / 1e12 / a.peak * 1e3:.0f} + "
          f"{n_snip * attn / 1e12 / a.peak * 1e3:.0f} ms at {a.peak} TFLOP/s")
    n_dec = 3 * n_snip
    print(f"VAE encoder: {enc['conv'] / 1e12:.3f} TFLOP of convolutions per {a.height}x{a.width} frame")
    print(f"VAE decoder: {dec['conv'] / 1e12:.3f} TFLOP of convolutions per frame (+{dec['linear'] / 1e9:.1f} GFLOP linear; "
          f"{dec['subpixel_saved'] / 1e12:.3f} TFLOP saved by the sub-pixel form of the three upsampling convolutions)")
    print(f"{a.frames} frames, dilations {a.dilations}: {n_snip} snippets -> {n_dec} decoded frames = "
          f"{n_dec * dec['conv'] / 1e15:.3f} PFLOP = {n_dec * dec['conv'] / 1e12 / a.peak * 1e3:.0f} ms at {a.peak} TFLOP/s; "
          f"encode {a.frames * enc['conv'] / 1e15:.3f} PFLOP = {a.frames * enc['conv'] / 1e12 / a.peak * 1e3:.0f} ms")


if __name__ == "__main__":
    main()
