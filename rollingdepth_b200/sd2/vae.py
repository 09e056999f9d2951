"""SD-architecture `AutoencoderKL` (encoder / decoder / quant convs) as used by RollingDepth.

Behavioural reference: diffusers/src/diffusers/models/autoencoders/{autoencoder_kl.py, vae.py:53-347} and
unets/unet_2d_blocks.py:UNetMidBlock2D / DownEncoderBlock2D / UpDecoderBlock2D.  The pipeline calls
`vae.encoder`, `vae.quant_conv`, `vae.post_quant_conv` and `vae.decoder` directly
(rollingdepth_pipeline.py:690-693, :731-732), so those four attributes are the surface.  Parameter names and
shapes match diffusers (new-style `to_q/to_k/to_v/to_out.0` attention keys).  Arithmetic stays on
cuDNN / cuBLAS / torch SDPA.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Sequence

import torch
import torch.nn as nn

from .layers import ANY, Attention, BiasConv2d, check_ignored_config, Downsample2D, GroupNormAct, ResnetBlock2D, Upsample2D

SD_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                     layers_per_block=2, norm_num_groups=32)


class _VaeMid(nn.Module):
    def __init__(self, c: int, groups: int, attention_head_dim: int):
        super().__init__()
        self.attentions = nn.ModuleList([
            Attention(c, None, heads=c // attention_head_dim, dim_head=attention_head_dim, bias=True,
                      norm_num_groups=groups, eps=1e-6, residual_connection=True, rescale_output_factor=1.0)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, groups, eps=1e-6), ResnetBlock2D(c, c, None, groups, eps=1e-6)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, eps=1e-6) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class _DecBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, eps=1e-6) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels: Sequence[int], layers_per_block, groups, double_z=True):
        super().__init__()
        self.conv_in = BiasConv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = block_out_channels[0]
        for i, c in enumerate(block_out_channels):
            cin, cout = cout, c
            self.down_blocks.append(_EncBlock(cin, cout, layers_per_block, groups, i != len(block_out_channels) - 1))
        self.mid_block = _VaeMid(block_out_channels[-1], groups, block_out_channels[-1])
        self.conv_norm_out = GroupNormAct(groups, block_out_channels[-1], eps=1e-6, act=True)
        self.conv_act = nn.SiLU()
        self.conv_out = BiasConv2d(block_out_channels[-1], 2 * out_channels if double_z else out_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_norm_out(x))      # conv_norm_out carries the SiLU (conv_act)


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels: Sequence[int], layers_per_block, groups):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = BiasConv2d(in_channels, rev[0], 3, padding=1)
        self.mid_block = _VaeMid(rev[0], groups, rev[0])
        self.up_blocks = nn.ModuleList()
        cout = rev[0]
        for i, c in enumerate(rev):
            cin, cout = cout, c
            self.up_blocks.append(_DecBlock(cin, cout, layers_per_block + 1, groups, i != len(rev) - 1))
        self.conv_norm_out = GroupNormAct(groups, rev[-1], eps=1e-6, act=True)
        self.conv_act = nn.SiLU()
        self.conv_out = BiasConv2d(rev[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_norm_out(x))      # conv_norm_out carries the SiLU (conv_act)


# diffusers AutoencoderKL config keys (autoencoder_kl.py:70-100) that are not constructor parameters here
_VAE_IGNORED_OK = dict(
    sample_size=ANY, act_fn=("silu",), down_block_types=ANY, up_block_types=ANY, force_upcast=ANY,
    latents_mean=(None,), latents_std=(None,), shift_factor=(None,), use_quant_conv=(True,), use_post_quant_conv=(True,),
    mid_block_add_attention=(True,))


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(64,), layers_per_block=1,
                 norm_num_groups=32, scaling_factor=0.18215, **unused):
        super().__init__()
        check_ignored_config("AutoencoderKL", unused, _VAE_IGNORED_OK)
        for key in ("down_block_types", "up_block_types"):
            for t in unused.get(key, ()):
                if t not in ("DownEncoderBlock2D", "UpDecoderBlock2D"):
                    raise NotImplementedError(f"AutoencoderKL: {key} entry {t!r} is not implemented")
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups, True)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = BiasConv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = BiasConv2d(latent_channels, latent_channels, 1)

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    @property
    def device(self):
        return self.quant_conv.weight.device
