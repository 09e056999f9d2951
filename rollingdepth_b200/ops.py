"""Functional wrappers over the C ABI (include/rd_b200.h) taking torch CUDA tensors.

PyTorch is used for device memory and streams only; all arithmetic happens in librd_b200.so.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib as L


def xattn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
    """softmax(scale * q k^T) v over [B, S, H, D] views (any batch/seq/head strides, D contiguous).

    Replaces F.scaled_dot_product_attention at attention_processor.py:2251-2253 of the reference;
    q/k/v are the `[B, S, H*D]` projections viewed as `[B, S, H, D]` -- no transpose copies.
    """
    L.require_cuda(q, k, v)
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise L.RdError("xattn_fwd expects [B, S, H, D] tensors")
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    if k.shape != (B, Sk, H, D) or v.shape != (B, Sk, H, D):
        raise L.RdError(f"shape mismatch q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)}")
    if not (q.dtype == k.dtype == v.dtype):
        raise L.RdError("q, k, v must share a dtype")
    if q.stride(3) != 1 or k.stride(3) != 1 or v.stride(3) != 1:
        raise L.RdError("head_dim axis must be contiguous")
    if out is None:
        out = torch.empty((B, Sq, H, D), dtype=q.dtype, device=q.device)
    elif tuple(out.shape) != (B, Sq, H, D) or out.dtype != q.dtype or out.device != q.device or out.stride(3) != 1:
        raise L.RdError(f"out must be a [B, Sq, H, D] = {(B, Sq, H, D)} tensor of q's dtype / device with a contiguous head_dim axis")
    scale = (1.0 / math.sqrt(D)) if scale is None else float(scale)
    L.call(
        "rd_xattn_fwd", q.device, L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(out), B, H, Sq, Sk, D,
        q.stride(0), q.stride(1), q.stride(2),
        k.stride(0), k.stride(1), k.stride(2),
        v.stride(0), v.stride(1), v.stride(2),
        out.stride(0), out.stride(1), out.stride(2),
        L.dtype_code(q.dtype), scale, int(flags), L.stream_ptr(q.device))
    L.count_launches(1)
    return out


_PTYPE = {"epsilon": 0, "sample": 1, "v_prediction": 2}


def ddim_step(model_output: torch.Tensor, sample: torch.Tensor, sqrt_alpha_t: float, sqrt_beta_t: float,
              sqrt_alpha_prev: float, sqrt_one_minus_alpha_prev: float, prediction_type: str,
              clip: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One DDIM step (eta = 0) with the reference's op-by-op rounding (scheduling_ddim.py:342-468)."""
    L.require_cuda(model_output, sample)
    if model_output.shape != sample.shape or model_output.dtype != sample.dtype:
        raise L.RdError("model_output and sample must match in shape and dtype")
    mo, sm = model_output.contiguous(), sample.contiguous()
    if out is None:
        out = torch.empty_like(sm)
    elif out.shape != sm.shape or out.dtype != sm.dtype or out.device != sm.device or not out.is_contiguous():
        raise L.RdError("out must be a contiguous tensor of the sample's shape, dtype and device")
    L.call("rd_ddim_step", sm.device, L.ptr(mo), L.ptr(sm), L.ptr(out), sm.numel(), sqrt_alpha_t, sqrt_beta_t,
           sqrt_alpha_prev, sqrt_one_minus_alpha_prev, _PTYPE[prediction_type], float(clip),
           L.dtype_code(sm.dtype), L.stream_ptr(sm.device))
    L.count_launches(1)
    return out


def ddim_add_noise(original: torch.Tensor, noise: torch.Tensor, sqrt_alpha_prod: float,
                   sqrt_one_minus_alpha_prod: float) -> torch.Tensor:
    """scheduling_ddim.py:471-495."""
    L.require_cuda(original, noise)
    x, n = original.contiguous(), noise.contiguous()
    out = torch.empty_like(x)
    L.call("rd_ddim_add_noise", x.device, L.ptr(x), L.ptr(n), L.ptr(out), x.numel(), sqrt_alpha_prod,
           sqrt_one_minus_alpha_prod, L.dtype_code(x.dtype), L.stream_ptr(x.device))
    L.count_launches(1)
    return out


def renorm_(depth: torch.Tensor, minmax: torch.Tensor) -> torch.Tensor:
    """In place d = ((d - mn) / (mx - mn)) * 2 - 1 (rollingdepth_pipeline.py:316-318)."""
    L.require_cuda(depth, minmax)
    if not depth.is_contiguous():
        raise L.RdError("renorm_ needs a contiguous tensor")
    L.call("rd_renorm", depth.device, L.ptr(depth), depth.numel(), L.ptr(minmax), L.dtype_code(depth.dtype),
           L.stream_ptr(depth.device))
    L.count_launches(1)
    return depth


def _nhwc_dims(x: torch.Tensor):
    if x.dim() == 4:
        N, C, H, W = x.shape
        if not x.is_contiguous(memory_format=torch.channels_last):
            raise L.RdError("expected a channels_last [N, C, H, W] tensor")
        return N, H * W, C
    if x.dim() == 3:
        if not x.is_contiguous():
            raise L.RdError("expected a contiguous [N, L, C] tensor")
        return x.shape[0], x.shape[1], x.shape[2]
    raise L.RdError("expected a 3-D [N, L, C] or 4-D channels_last tensor")


def _opt(t: Optional[torch.Tensor], dtype) -> Optional[torch.Tensor]:
    return None if t is None else t.to(dtype).contiguous()


def group_norm_nhwc(x: torch.Tensor, num_groups: int, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                    eps: float, silu: bool, pre_add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """act(GroupNorm(x + pre_add[n, c])) on a channels-last activation ([N, C, H, W] in torch.channels_last
    memory format or [N, L, C] contiguous); same shape / memory format out."""
    L.require_cuda(x)
    N, HW, Cc = _nhwc_dims(x)
    y = torch.empty_like(x)          # preserve_format keeps channels_last
    lib = L.lib()
    ws_bytes = lib.rd_groupnorm_workspace_bytes(N, num_groups)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    g, b, pre = _opt(weight, x.dtype), _opt(bias, x.dtype), _opt(pre_add, x.dtype)
    if pre is not None and tuple(pre.shape) != (N, Cc):
        raise L.RdError(f"pre_add must be [{N}, {Cc}], got {tuple(pre.shape)}")
    P = lambda t: L.ptr(t) if t is not None else None
    L.call("rd_groupnorm_nhwc", x.device, L.ptr(x), L.ptr(y), P(g), P(b), P(pre), N, HW, Cc, int(num_groups), float(eps),
           1 if silu else 0, L.dtype_code(x.dtype), L.ptr(ws), ws_bytes, L.stream_ptr(x.device))
    L.count_launches(2)
    return y


def bias_residual_nhwc(x: torch.Tensor, bias: Optional[torch.Tensor] = None, vec: Optional[torch.Tensor] = None,
                       residual: Optional[torch.Tensor] = None, inplace: bool = True) -> torch.Tensor:
    """x + bias[c] + vec[n, c] + residual in one pass (channels-last)."""
    L.require_cuda(x)
    N, HW, Cc = _nhwc_dims(x)
    if residual is not None:
        if residual.shape != x.shape or residual.dtype != x.dtype:
            raise L.RdError("residual must match x")
        _nhwc_dims(residual)
    y = x if inplace else torch.empty_like(x)
    b, v = _opt(bias, x.dtype), _opt(vec, x.dtype)
    P = lambda t: L.ptr(t) if t is not None else None
    L.call("rd_bias_residual_nhwc", x.device, L.ptr(x), P(b), P(v), P(residual), L.ptr(y), N, HW, Cc,
           L.dtype_code(x.dtype), L.stream_ptr(x.device))
    L.count_launches(1)
    return y


def upsample_nearest_nhwc(x: torch.Tensor, size) -> torch.Tensor:
    """F.interpolate(x, size=size, mode="nearest") for a channels_last [N, C, H, W] tensor."""
    L.require_cuda(x)
    if x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last):
        raise L.RdError("expected a channels_last [N, C, H, W] tensor")
    N, Cc, H, W = x.shape
    Ho, Wo = int(size[0]), int(size[1])
    y = torch.empty((N, Cc, Ho, Wo), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    L.call("rd_upsample_nearest_nhwc", x.device, L.ptr(x), L.ptr(y), N, H, W, Ho, Wo, Cc, L.dtype_code(x.dtype),
           L.stream_ptr(x.device))
    L.count_launches(1)
    return y


def phase_shuffle_bias_nhwc(t: torch.Tensor, bias, H: int, W: int) -> torch.Tensor:
    """y[n, c, 2i + a, 2j + b] = t[n, (2a + b) C + c, i + a, j + b] + bias[c] for a channels_last t of shape
    [N, 4 C, H + 1, W + 1] (the four phases of a sub-pixel Upsample2D convolution) -> channels_last [N, C, 2 H, 2 W]."""
    L.require_cuda(t)
    if t.dim() != 4 or not t.is_contiguous(memory_format=torch.channels_last):
        raise L.RdError("expected a channels_last [N, 4C, H+1, W+1] tensor")
    N, C4, Hp, Wp = t.shape
    if C4 % 4 != 0 or Hp != H + 1 or Wp != W + 1:
        raise L.RdError(f"phase tensor {tuple(t.shape)} does not match H={H}, W={W}")
    Cc = C4 // 4
    if bias is not None and (bias.dtype != t.dtype or bias.numel() != Cc or not bias.is_contiguous()):
        raise L.RdError("bias must be a contiguous [C] tensor of the activation dtype")
    y = torch.empty((N, Cc, 2 * H, 2 * W), dtype=t.dtype, device=t.device, memory_format=torch.channels_last)
    L.call("rd_phase_shuffle_bias_nhwc", t.device, L.ptr(t), L.ptr(bias) if bias is not None else None, L.ptr(y),
           N, H, W, Cc, L.dtype_code(t.dtype), L.stream_ptr(t.device))
    L.count_launches(1)
    return y


def geglu(proj: torch.Tensor) -> torch.Tensor:
    """proj[..., :D] * gelu(proj[..., D:]) for a contiguous [..., 2 D] projection (one fused pass)."""
    L.require_cuda(proj)
    if not proj.is_contiguous():
        proj = proj.contiguous()
    D = proj.shape[-1] // 2
    M = proj.numel() // (2 * D)
    y = torch.empty(*proj.shape[:-1], D, dtype=proj.dtype, device=proj.device)
    L.call("rd_geglu", proj.device, L.ptr(proj), L.ptr(y), M, D, L.dtype_code(proj.dtype), L.stream_ptr(proj.device))
    L.count_launches(1)
    return y


SOFTMAX_ROWS_MAX_COLS = 10240


def softmax_rows(logits: torch.Tensor, scale: float, out_dtype: torch.dtype) -> torch.Tensor:
    """softmax(scale * logits, dim=-1).to(out_dtype) for contiguous fp32 logits [..., cols] in one pass
    (cols % 4 == 0, cols <= SOFTMAX_ROWS_MAX_COLS)."""
    L.require_cuda(logits)
    if logits.dtype != torch.float32 or not logits.is_contiguous():
        raise L.RdError("softmax_rows expects contiguous float32 logits")
    cols = logits.shape[-1]
    rows = logits.numel() // cols
    out = torch.empty(logits.shape, dtype=out_dtype, device=logits.device)
    L.call("rd_softmax_rows", logits.device, L.ptr(logits), L.ptr(out), rows, cols, float(scale), L.dtype_code(out_dtype),
           L.stream_ptr(logits.device))
    L.count_launches(1)
    return out


def layer_norm_supported(x: torch.Tensor, C: int) -> bool:
    vn = 16 // x.element_size()
    return x.is_cuda and x.dtype in _LN_DTYPES and C % vn == 0 and C <= 32 * 8 * vn


_LN_DTYPES = (torch.float16, torch.bfloat16, torch.float32)


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
    """F.layer_norm(x, (C,), weight, bias, eps) over the last axis of a contiguous tensor in one pass."""
    L.require_cuda(x)
    if not x.is_contiguous():
        x = x.contiguous()
    Cc = x.shape[-1]
    M = x.numel() // Cc
    y = torch.empty_like(x)
    g, b = _opt(weight, x.dtype), _opt(bias, x.dtype)
    L.call("rd_layernorm", x.device, L.ptr(x), L.ptr(y), L.ptr(g), L.ptr(b), M, Cc, float(eps), L.dtype_code(x.dtype),
           L.stream_ptr(x.device))
    L.count_launches(1)
    return y
