"""`B200XFrameAttnProcessor` -- a diffusers AttnProcessor that runs RollingDepth's cross-frame
self-attention on the hand-written sm_100a kernel.

Interface and semantics follow the reference's default processor, `AttnProcessor2_0.__call__`
(diffusers/src/diffusers/models/attention_processor.py:2181-2276, the "Modified in RollingDepth"
version) and its xformers twin (:1961-2060):

  * `num_view` is named in the signature -- `Attention.forward` (:483-492) silently drops kwargs the
    processor does not declare, which would degrade to per-frame attention;
  * self-attention with `num_view = n`:  "(b n) hw c -> b (n hw) c" before the projections and the
    inverse after `to_out` (:2208-2211, :2263-2266) -- both are views, frames are already contiguous;
  * text cross-attention: the reference merges all frames into one query sequence because the batch
    is taken from the context (:2213-2215); a context of batch 1 or b is accepted here;
  * q/k/v stay in the `[B, S, heads*64]` layout the Linear layers produce and are handed to the kernel
    as strided `[B, S, heads, 64]` views: no head-major copies (the xformers path makes three, :2032-2034).

Install with `unet.set_attn_processor(B200XFrameAttnProcessor())` (unets/unet_2d_condition.py:721-753)
or `Attention.set_processor` (:422-440).  The object is stateless and shared by all layers.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import ops


class B200XFrameAttnProcessor:
    def __init__(self, kernel_flags: int = 0):
        self.kernel_flags = int(kernel_flags)
        # bench.py sets this to a list to collect (shape, start_event, end_event) per kernel launch
        self.events = None

    def __call__(
        self,
        attn,
        hidden_states: torch.Tensor,
        encoder_hidden_states: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        temb: Optional[torch.Tensor] = None,
        num_view: int = None,
        *args,
        **kwargs,
    ) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("B200XFrameAttnProcessor: attention masks are not used on the RollingDepth path")
        L.require_cuda(hidden_states)       # no CPU fallback

        residual = hidden_states
        if getattr(attn, "spatial_norm", None) is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)

        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            batch_size, channel, height, width = hidden_states.shape
            hidden_states = hidden_states.reshape(batch_size, channel, height * width).transpose(1, 2)

        if num_view is not None:                                     # "(b n) hw c -> b (n hw) c"
            bn, hw, c = hidden_states.shape
            if bn % num_view != 0:
                raise ValueError(f"batch {bn} is not a multiple of num_view {num_view}")
            hidden_states = hidden_states.reshape(bn // num_view, num_view * hw, c)

        if getattr(attn, "group_norm", None) is not None:
            gn = attn.group_norm
            if hasattr(gn, "forward_nlc"):      # our GroupNormAct: NHWC kernel, no transposes
                hidden_states = gn.forward_nlc(hidden_states)
            else:
                hidden_states = gn(hidden_states.transpose(1, 2)).transpose(1, 2)

        query = attn.to_q(hidden_states)

        if encoder_hidden_states is None:
            context = hidden_states
        else:
            context = encoder_hidden_states
            if getattr(attn, "norm_cross", None):
                context = attn.norm_encoder_hidden_states(context)
            if context.shape[0] != hidden_states.shape[0]:
                if context.shape[0] == 1 and num_view is not None and hidden_states.shape[0] > 1:
                    context = context.expand(hidden_states.shape[0], -1, -1)
                elif context.shape[0] == 1 and num_view is None:
                    # per-frame path with a shared context: reference xformers processor repeats it (:1997-2001)
                    context = context.expand(hidden_states.shape[0], -1, -1)
                else:
                    raise ValueError(
                        f"context batch {context.shape[0]} does not match hidden batch {hidden_states.shape[0]}")
        batch = hidden_states.shape[0]

        key = attn.to_k(context)
        value = attn.to_v(context)

        heads = attn.heads
        head_dim = key.shape[-1] // heads
        q4 = query.view(batch, -1, heads, head_dim)
        k4 = key.view(batch, -1, heads, head_dim)
        v4 = value.view(batch, -1, heads, head_dim)
        if getattr(attn, "norm_q", None) is not None:
            q4 = attn.norm_q(q4)
        if getattr(attn, "norm_k", None) is not None:
            k4 = attn.norm_k(k4)

        # reference: F.scaled_dot_product_attention default scale 1/sqrt(head_dim) (:2251)
        ev = self.events
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        out = ops.xattn_fwd(q4, k4, v4, scale=None, flags=self.kernel_flags)
        if ev is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            ev.append(((batch, q4.shape[1], k4.shape[1], heads), e0, e1))
        hidden_states = out.reshape(batch, -1, heads * head_dim).to(query.dtype)

        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)

        if num_view is not None:                                     # "b (n hw) c -> (b n) hw c"
            b, s, c = hidden_states.shape
            hidden_states = hidden_states.reshape(b * num_view, s // num_view, c)

        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(batch_size, channel, height, width)

        if getattr(attn, "residual_connection", False):
            hidden_states = hidden_states + residual

        rescale = getattr(attn, "rescale_output_factor", 1.0)
        if rescale != 1.0:
            hidden_states = hidden_states / rescale
        return hidden_states
