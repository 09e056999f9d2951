"""Oracle for the cross-frame self-attention layer (TEST INFRASTRUCTURE).

Restates diffusers/src/diffusers/models/attention_processor.py:2181-2276
(`AttnProcessor2_0.__call__`, the "Modified in RollingDepth" default processor)
and :1961-2060 (`XFormersAttnProcessor`, same maths).  The fused library call at
:2251 (torch SDPA; torch is a third-party dependency of the reference, pinned
torch==2.4.1 in requirements.txt:1) is restated from its published definition
softmax(Q K^T * scale) V with scale = 1/sqrt(head_dim), evaluated in fp32 or
fp64 and chunked over queries so the score matrix is never fully materialised.
Pinned by tests/golden/xattn_*.npz produced from the reference processor.
"""
from __future__ import annotations

import math
from typing import Optional

import torch


def sdpa_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None,
             acc_dtype: torch.dtype = torch.float32, chunk: int = 1024) -> torch.Tensor:
    """q [B,H,Sq,D], k/v [B,H,Sk,D] -> [B,H,Sq,D] in q.dtype; math in acc_dtype."""
    B, H, Sq, D = q.shape
    scale = (1.0 / math.sqrt(D)) if scale is None else scale
    qf, kf, vf = q.to(acc_dtype), k.to(acc_dtype), v.to(acc_dtype)
    out = torch.empty(B, H, Sq, v.shape[-1], dtype=acc_dtype, device=q.device)
    for s0 in range(0, Sq, chunk):
        sc = torch.matmul(qf[:, :, s0:s0 + chunk], kf.transpose(-1, -2)) * scale
        p = torch.softmax(sc, dim=-1)
        out[:, :, s0:s0 + chunk] = torch.matmul(p, vf)
    return out.to(q.dtype)


def xframe_processor_ref(
    hidden_states: torch.Tensor,          # [(b n), hw, C]
    to_q: torch.Tensor, to_k: torch.Tensor, to_v: torch.Tensor,   # [C_inner, C_in] bias-free (:224-229)
    to_out_w: torch.Tensor, to_out_b: Optional[torch.Tensor],
    heads: int,
    encoder_hidden_states: Optional[torch.Tensor] = None,
    num_view: Optional[int] = None,
    residual_connection: bool = False,
    rescale_output_factor: float = 1.0,
    acc_dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """attention_processor.py:2198-2276 for 3-D inputs, no mask / norms (SD2 UNet)."""
    residual = hidden_states
    x = hidden_states
    if num_view is not None:                                   # :2208-2211  "(b n) hw c -> b (n hw) c"
        bn, hw, c = x.shape
        x = x.reshape(bn // num_view, num_view * hw, c)
    ctx = x if encoder_hidden_states is None else encoder_hidden_states
    batch = ctx.shape[0]                                       # :2213-2215 batch comes from ctx!
    q = torch.nn.functional.linear(x, to_q)                    # :2226
    k = torch.nn.functional.linear(ctx, to_k)                  # :2233
    v = torch.nn.functional.linear(ctx, to_v)
    d = k.shape[-1] // heads
    q = q.view(batch, -1, heads, d).transpose(1, 2)            # :2239-2242
    k = k.view(batch, -1, heads, d).transpose(1, 2)
    v = v.view(batch, -1, heads, d).transpose(1, 2)
    o = sdpa_ref(q, k, v, acc_dtype=acc_dtype)                 # :2251-2253
    o = o.transpose(1, 2).reshape(batch, -1, heads * d).to(q.dtype)   # :2255-2256
    o = torch.nn.functional.linear(o, to_out_w, to_out_b)      # :2259 (dropout p=0)
    if num_view is not None:                                   # :2263-2266  "b (n hw) c -> (b n) hw c"
        b, s, c = o.shape
        o = o.reshape(b * num_view, s // num_view, c)
    if residual_connection:
        o = o + residual
    return o / rescale_output_factor
