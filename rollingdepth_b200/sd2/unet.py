"""SD2-architecture `UNet2DConditionModel` with RollingDepth's `num_view` plumbing.

Behavioural reference: diffusers/src/diffusers/models/unets/unet_2d_condition.py:1039-1324 (forward),
unet_2d_blocks.py:847-908 / 1250-1327 / 2500-2591 (mid / down / up cross-attention blocks).  The only
RollingDepth change there is threading `num_view` from `forward` down to every attention processor;
that is reproduced here.  Parameter names/shapes equal diffusers' so checkpoints load unchanged.
Convolutions / GEMMs / norms run on cuDNN / cuBLAS through torch; the cross-frame
self-attention is dispatched to whatever processor is installed (`set_attn_processor`).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from .layers import (ANY, check_ignored_config, Attention, BiasConv2d, Downsample2D, GroupNormAct, ResnetBlock2D, TimestepEmbedding, Transformer2DModel, Upsample2D,
                     timestep_embedding)

SD2_UNET_CONFIG = dict(
    in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, attention_head_dim=(5, 10, 20, 20),
    cross_attention_dim=1024, use_linear_projection=True, norm_num_groups=32,
)


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, layers, groups, add_down, attn_cfg):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups) for i in range(layers)])
        if attn_cfg is not None:
            heads, cross, linear = attn_cfg
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, cout // heads, cout, 1, cross, groups, linear) for _ in range(layers)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=1)]) if add_down else None

    def forward(self, x, temb, ctx, num_view):
        outs = ()
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, encoder_hidden_states=ctx, num_view=num_view)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class _MidBlock(nn.Module):
    def __init__(self, c, temb, groups, heads, cross, linear):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups), ResnetBlock2D(c, c, temb, groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, 1, cross, groups, linear)])

    def forward(self, x, temb, ctx, num_view):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, encoder_hidden_states=ctx, num_view=num_view)
        return self.resnets[1](x, temb)


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, temb, layers, groups, add_up, attn_cfg):
        super().__init__()
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = cprev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb, groups))
        self.resnets = nn.ModuleList(res)
        if attn_cfg is not None:
            heads, cross, linear = attn_cfg
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, cout // heads, cout, 1, cross, groups, linear) for _ in range(layers)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips: Tuple[torch.Tensor, ...], temb, ctx, num_view, upsample_size=None):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips[-1]], dim=1)
            skips = skips[:-1]
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, encoder_hidden_states=ctx, num_view=num_view)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x, upsample_size)
        return x


# diffusers UNet2DConditionModel config keys (unet_2d_condition.py:170-224) that are not constructor parameters here,
# with the only values this implementation covers.  `upcast_attention` is accepted either way: both reference
# processors ignore it on their fused paths, and so does ours.
_UNET_IGNORED_OK = dict(
    sample_size=ANY, center_input_sample=(False,), act_fn=("silu",), norm_eps=(1e-5,), downsample_padding=(1,),
    mid_block_type=("UNetMidBlock2DCrossAttn",), mid_block_scale_factor=(1, 1.0), mid_block_only_cross_attention=(None,),
    only_cross_attention=(False,), dual_cross_attention=(False,), dropout=(0.0, 0), transformer_layers_per_block=(1,),
    reverse_transformer_layers_per_block=(None,), encoder_hid_dim=(None,), encoder_hid_dim_type=(None,),
    class_embed_type=(None,), addition_embed_type=(None,), addition_time_embed_dim=(None,), num_class_embeds=(None,),
    upcast_attention=ANY, resnet_time_scale_shift=("default",), resnet_skip_time_act=(False,),
    resnet_out_scale_factor=(1.0, 1), time_embedding_type=("positional",), time_embedding_dim=(None,),
    time_embedding_act_fn=(None,), timestep_post_act=(None,), time_cond_proj_dim=(None,), conv_in_kernel=(3,),
    conv_out_kernel=(3,), projection_class_embeddings_input_dim=(None,), attention_type=("default",),
    class_embeddings_concat=(False,), cross_attention_norm=(None,), addition_embed_type_num_heads=ANY,
    up_block_types=ANY, down_block_types=ANY)


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, down_block_types=SD2_UNET_CONFIG["down_block_types"],
                 up_block_types=SD2_UNET_CONFIG["up_block_types"], block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block=2, attention_head_dim: Union[int, Sequence[int]] = 8, num_attention_heads=None,
                 cross_attention_dim=1280, use_linear_projection=False, norm_num_groups=32, flip_sin_to_cos=True,
                 freq_shift=0, **unused):
        super().__init__()
        check_ignored_config("UNet2DConditionModel", unused, _UNET_IGNORED_OK)
        self.config = SimpleNamespace(
            in_channels=in_channels, out_channels=out_channels, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
            layers_per_block=layers_per_block, attention_head_dim=attention_head_dim,
            cross_attention_dim=cross_attention_dim, use_linear_projection=use_linear_projection,
            norm_num_groups=norm_num_groups, flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift)
        # diffusers quirk (unet_2d_condition.py:242): `attention_head_dim` is the head COUNT
        heads = num_attention_heads or attention_head_dim
        nb = len(block_out_channels)
        heads = (heads,) * nb if isinstance(heads, int) else tuple(heads)
        temb = block_out_channels[0] * 4
        self.conv_in = BiasConv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], temb)

        self.down_blocks = nn.ModuleList()
        cout = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            cin, cout = cout, block_out_channels[i]
            cfg = (heads[i], cross_attention_dim, use_linear_projection) if t == "CrossAttnDownBlock2D" else None
            if t not in ("CrossAttnDownBlock2D", "DownBlock2D"):
                raise NotImplementedError(t)
            self.down_blocks.append(_DownBlock(cin, cout, temb, layers_per_block, norm_num_groups, i != nb - 1, cfg))

        self.mid_block = _MidBlock(block_out_channels[-1], temb, norm_num_groups, heads[-1], cross_attention_dim,
                                   use_linear_projection)

        self.up_blocks = nn.ModuleList()
        rev, rheads = list(reversed(block_out_channels)), list(reversed(heads))
        cout = rev[0]
        for i, t in enumerate(up_block_types):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, nb - 1)]
            cfg = (rheads[i], cross_attention_dim, use_linear_projection) if t == "CrossAttnUpBlock2D" else None
            if t not in ("CrossAttnUpBlock2D", "UpBlock2D"):
                raise NotImplementedError(t)
            self.up_blocks.append(_UpBlock(cin, cout, cprev, temb, layers_per_block + 1, norm_num_groups, i != nb - 1, cfg))

        self.conv_norm_out = GroupNormAct(norm_num_groups, block_out_channels[0], eps=1e-5, act=True)
        self.conv_act = nn.SiLU()
        self.conv_out = BiasConv2d(block_out_channels[0], out_channels, 3, padding=1)

    # ------------------------------------------------------------------ processor registry (unet_2d_condition.py:697-753)
    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{name}.processor": m.processor for name, m in self.named_modules() if isinstance(m, Attention)}

    def set_attn_processor(self, processor) -> None:
        mods = {name: m for name, m in self.named_modules() if isinstance(m, Attention)}
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does "
                                 f"not match the number of attention layers: {len(mods)}.")
            for name, m in mods.items():
                m.set_processor(processor[f"{name}.processor"])
        else:
            for m in mods.values():
                m.set_processor(processor)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    # ------------------------------------------------------------------ forward
    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor, num_view: Optional[int] = None,
                return_dict: bool = True):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = timestep_embedding(timesteps, self.config.block_out_channels[0], self.config.flip_sin_to_cos,
                                   self.config.freq_shift).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)

        x = self.conv_in(sample)
        skips = (x,)
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states, num_view)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states, num_view)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            take, skips = skips[-n:], skips[:-n]
            # diffusers hands the next skip's size to the upsampler when the latent size is not a multiple of
            # the total stride (unet_2d_condition.py:1283-1287); for exact multiples this equals scale_factor=2
            size = skips[-1].shape[-2:] if (blk.upsamplers is not None and len(skips) > 0) else None
            x = blk(x, take, emb, encoder_hidden_states, num_view, size)
        x = self.conv_out(self.conv_norm_out(x))        # conv_norm_out carries the SiLU (conv_act)
        if not return_dict:
            return (x,)
        return SimpleNamespace(sample=x)
