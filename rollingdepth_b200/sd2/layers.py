"""Building blocks of the SD2-architecture UNet / VAE that RollingDepth runs.

These are the *callers* of the hot kernel: plain PyTorch modules whose
convolutions, GEMMs and norms stay on cuDNN / cuBLAS, and whose `Attention` module exposes the same
plug-in surface as diffusers' (`processor(attn, hidden_states, encoder_hidden_states=..., num_view=...)`,
`set_processor`, `to_q/to_k/to_v/to_out`, `heads`, `scale`), so that `B200XFrameAttnProcessor` is a
drop-in on either.  Parameter names and shapes match diffusers 0.30 so a `prs-eth/rollingdepth-v1-0`
(or any SD2-layout) state dict loads with `strict=True`.

Reference for behaviour: diffusers/src/diffusers/models/{attention_processor.py:45-716, attention.py:421-548,
transformers/transformer_2d.py:327-483, resnet.py:200-376, downsampling.py, upsampling.py, embeddings.py}.
"""
from __future__ import annotations

import inspect
import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- config hygiene
ANY = object()


def check_ignored_config(kind: str, ignored: dict, implemented: dict) -> None:
    """Constructor kwargs of a diffusers `config.json` that these modules do not take as parameters: each must have the
    one value this implementation hard-codes (or be listed as irrelevant with `ANY`); anything else -- a key we do not
    know, or a known key with another value -- would load with strict=True and silently compute something different,
    so it raises instead."""
    for k, v in ignored.items():
        if k.startswith("_"):                       # _class_name, _diffusers_version, _name_or_path
            continue
        if k not in implemented:
            raise NotImplementedError(f"{kind}: config key {k!r} (= {v!r}) is not supported by rollingdepth_b200.sd2")
        ok = implemented[k]
        if ok is ANY:
            continue
        v_cmp = tuple(v) if isinstance(v, list) else v
        if not any(v_cmp == o for o in ok):
            raise NotImplementedError(f"{kind}: {k} = {v!r} is not implemented (supported: {list(ok)})")


# --------------------------------------------------------------------------- time embedding
def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, freq_shift: float = 0.0,
                       max_period: int = 10000) -> torch.Tensor:
    """Sinusoidal embedding (embeddings.py get_timestep_embedding), fp32 math."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------- convolution with a fused bias pass
def _is_cl(x: torch.Tensor) -> bool:
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)


class BiasConv2d(nn.Conv2d):
    """`nn.Conv2d` (same parameters / state-dict keys).  With `fused = True` on CUDA the convolution runs
    bias-free on cuDNN (NHWC) and the bias is added by one vectorised in-place pass of our own kernel
    (rd_bias_residual_nhwc) instead of torch's broadcast add, which is not vectorised for NHWC tensors."""

    fused = False

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.fused and x.is_cuda and self.bias is not None and self.out_channels % (16 // x.element_size()) == 0:
            y = F.conv2d(x, self.weight, None, self.stride, self.padding, self.dilation, self.groups)
            if _is_cl(y):
                from .. import ops
                return ops.bias_residual_nhwc(y, bias=self.bias)
            return y + self.bias.view(1, -1, 1, 1)
        return super().forward(x)


# --------------------------------------------------------------------------- GroupNorm (+SiLU)
class GroupNormAct(nn.GroupNorm):
    """`nn.GroupNorm` (same parameters / state-dict keys) with an optional trailing SiLU.

    With `fused = True` (set by the B200 pipeline) CUDA channels-last inputs go to the hand-written NHWC
    kernel pair (C ABI rd_groupnorm_nhwc): statistics + normalise + SiLU in 2 reads and 1 write, no layout
    round trips.  Otherwise it is exactly F.group_norm followed by F.silu -- the fp32 reference of that kernel.
    """

    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5, act: bool = False):
        super().__init__(num_groups, num_channels, eps=eps, affine=True)
        self.act = act
        self.fused = False

    def _fusable(self, x: torch.Tensor) -> bool:
        return (self.fused and x.is_cuda and x.dtype in (torch.float16, torch.bfloat16, torch.float32)
                and self.num_channels % (16 // x.element_size()) == 0 and self.num_channels // (16 // x.element_size()) <= 512)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 4 and self._fusable(x) and x.is_contiguous(memory_format=torch.channels_last):
            from .. import ops
            return ops.group_norm_nhwc(x, self.num_groups, self.weight, self.bias, self.eps, self.act)
        y = F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
        return F.silu(y) if self.act else y

    def forward_nlc(self, x: torch.Tensor) -> torch.Tensor:
        """[N, L, C] in, [N, L, C] out (the layout attention processors hold)."""
        if self._fusable(x) and x.is_contiguous():
            from .. import ops
            return ops.group_norm_nhwc(x, self.num_groups, self.weight, self.bias, self.eps, self.act)
        y = F.group_norm(x.transpose(1, 2), self.num_groups, self.weight, self.bias, self.eps).transpose(1, 2)
        return F.silu(y) if self.act else y


def _to_nlc(x4: torch.Tensor) -> torch.Tensor:
    """[B, C, H, W] -> [B, H*W, C] without a copy for either memory format where possible."""
    b, c, h, w = x4.shape
    if x4.is_contiguous(memory_format=torch.channels_last) and not x4.is_contiguous():
        return x4.permute(0, 2, 3, 1).reshape(b, h * w, c)
    return x4.reshape(b, c, h * w).transpose(1, 2)


def _from_nlc(x3: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    b, c, h, w = like.shape
    if like.is_contiguous(memory_format=torch.channels_last) and not like.is_contiguous():
        return x3.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return x3.transpose(-1, -2).reshape(b, c, h, w)


def group_norm_nlc(attn, hidden_states: torch.Tensor) -> torch.Tensor:
    """attn.group_norm on a [B, L, C] tensor (attention_processor.py:2223-2224)."""
    gn = attn.group_norm
    if hasattr(gn, "forward_nlc"):
        return gn.forward_nlc(hidden_states)
    return gn(hidden_states.transpose(1, 2)).transpose(1, 2)


# --------------------------------------------------------------------------- attention (plug-in surface)
_GEMM_ATTENTION_OK = None      # torch.bmm(out_dtype=fp32) availability, probed once


def _use_gemm_attention(q: torch.Tensor) -> bool:
    """The VAE mid-block attention has ONE head of dim 512: torch's SDPA has no flash/cuDNN kernel beyond 256 and falls
    back to the sm80 memory-efficient kernel (175 TFLOP/s, 5.6 % of a `fast` step).  Two cuBLAS batched GEMMs with fp32
    logits + an fp32 softmax are ~2.5x faster and at least as accurate (P is rounded to 16 bits exactly like a fused
    kernel does)."""
    if not (_GEMM_ATTENTION_OK is not False and q.is_cuda and q.dtype in (torch.float16, torch.bfloat16)
            and q.shape[1] == 1 and q.shape[-1] > 256):
        return False
    # the fp32 logits are materialised ([B, S, S], twice with the softmax output): keep them below 24 GB
    return q.shape[0] * q.shape[2] * q.shape[2] * 4 <= (24 << 30)


def _wide_head_attention(q, k, v):
    global _GEMM_ATTENTION_OK
    q2, k2, v2 = q[:, 0], k[:, 0], v[:, 0]                # [B, S, d]
    try:
        logits = torch.bmm(q2, k2.transpose(1, 2), out_dtype=torch.float32)       # [B, Sq, Sk] fp32
        _GEMM_ATTENTION_OK = True
    except (TypeError, RuntimeError):
        _GEMM_ATTENTION_OK = False
        return None
    scale = q.shape[-1] ** -0.5
    from .. import ops
    if logits.shape[-1] % 4 == 0 and logits.shape[-1] <= ops.SOFTMAX_ROWS_MAX_COLS:
        p = ops.softmax_rows(logits, scale, q.dtype)      # scale, softmax and cast in one pass over the fp32 logits
    else:
        p = torch.softmax(logits.mul_(scale), dim=-1).to(q.dtype)
    return torch.bmm(p, v2)[:, None]                      # [B, 1, Sq, d]


class TorchSDPAProcessor:
    """Library attention (torch SDPA) with the reference's `num_view` semantics; used for the VAE
    mid-block (1 head, head_dim 512 -- SURVEY K4, left on the library) and selectable for A/B runs."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 num_view: int = None, *args, **kwargs):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            hidden_states = _to_nlc(hidden_states)
        if num_view is not None:
            bn, hw, c_ = hidden_states.shape
            hidden_states = hidden_states.reshape(bn // num_view, num_view * hw, c_)
        if attn.group_norm is not None:
            hidden_states = group_norm_nlc(attn, hidden_states)
        query = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        if ctx.shape[0] != hidden_states.shape[0]:
            ctx = ctx.expand(hidden_states.shape[0], -1, -1)
        key, value = attn.to_k(ctx), attn.to_v(ctx)
        batch = hidden_states.shape[0]
        d = key.shape[-1] // attn.heads
        q = query.view(batch, -1, attn.heads, d).transpose(1, 2)
        k = key.view(batch, -1, attn.heads, d).transpose(1, 2)
        v = value.view(batch, -1, attn.heads, d).transpose(1, 2)
        o = _wide_head_attention(q, k, v) if _use_gemm_attention(q) else None
        if o is None:
            o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(batch, -1, attn.heads * d).to(query.dtype)
        o = attn.to_out[1](attn.to_out[0](o))
        if num_view is not None:
            b_, s_, c_ = o.shape
            o = o.reshape(b_ * num_view, s_ // num_view, c_)
        if input_ndim == 4:
            o = _from_nlc(o, residual)
        if attn.residual_connection:
            o = o + residual
        return o / attn.rescale_output_factor


class Attention(nn.Module):
    """Same attribute surface as diffusers' Attention (attention_processor.py:45-268)."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 bias: bool = False, out_bias: bool = True, norm_num_groups: Optional[int] = None, eps: float = 1e-5,
                 residual_connection: bool = False, rescale_output_factor: float = 1.0, processor=None):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.query_dim = query_dim
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.residual_connection = residual_connection
        self.rescale_output_factor = rescale_output_factor
        self.group_norm = GroupNormAct(norm_num_groups, query_dim, eps=eps) if norm_num_groups else None
        self.spatial_norm = None
        self.norm_q = self.norm_k = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.processor = None
        self.set_processor(processor if processor is not None else TorchSDPAProcessor())

    def set_processor(self, processor) -> None:
        self.processor = processor
        self._proc_params = set(inspect.signature(processor.__call__).parameters.keys())

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        # diffusers filters kwargs by the processor's signature (attention_processor.py:483-492); a processor
        # that does not declare `num_view` would silently fall back to per-frame attention -- refuse instead.
        unknown = [k for k in cross_attention_kwargs if k not in self._proc_params]
        if "num_view" in unknown and cross_attention_kwargs["num_view"] is not None:
            raise TypeError(f"{type(self.processor).__name__} does not accept `num_view`")
        kw = {k: v for k, v in cross_attention_kwargs.items() if k in self._proc_params}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class LayerNormFused(nn.LayerNorm):
    """`nn.LayerNorm` (same parameters / state-dict keys).  With `fused = True` (set by the B200 pipeline) CUDA inputs go to the
    one-pass row kernel rd_layernorm (torch's vectorised LayerNorm reaches ~0.3 of the HBM bandwidth on the [B*S, 320..1280]
    activations of the transformer blocks and was 3 % of a `fast` step); otherwise exactly F.layer_norm."""

    fused = False

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.fused and self.elementwise_affine and self.bias is not None and len(self.normalized_shape) == 1:
            from .. import ops
            if ops.layer_norm_supported(x, self.normalized_shape[0]):
                return ops.layer_norm(x, self.weight, self.bias, self.eps)
        return super().forward(x)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    fused = True        # one rd_geglu pass on CUDA instead of gelu + mul over the two strided halves

    def forward(self, x):
        proj = self.proj(x)
        if self.fused and proj.is_cuda and proj.shape[-1] % 16 == 0:
            from .. import ops
            return ops.geglu(proj)
        h, gate = proj.chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    """attention.py:421-548 (layer_norm variant): LN -> attn1 (self, cross-frame) -> LN -> attn2 (text) -> LN -> GEGLU FF.
    `num_view` goes to attn1 AND attn2, as in the reference (:467-475, :508-516)."""

    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int]):
        super().__init__()
        self.norm1 = LayerNormFused(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = LayerNormFused(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = LayerNormFused(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None, num_view=None):
        x = self.attn1(self.norm1(x), encoder_hidden_states=None, num_view=num_view) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states, num_view=num_view) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    """transformers/transformer_2d.py:327-483, continuous-input path."""

    def __init__(self, heads: int, dim_head: int, in_channels: int, num_layers: int = 1,
                 cross_attention_dim: Optional[int] = None, norm_num_groups: int = 32, use_linear_projection: bool = True):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = GroupNormAct(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear_projection else BiasConv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels) if use_linear_projection else BiasConv2d(inner, in_channels, 1)

    def forward(self, x, encoder_hidden_states=None, num_view=None):
        b, c, h, w = x.shape
        residual = x
        x = self.norm(x)
        if self.use_linear_projection:
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
            x = self.proj_in(x)
        else:
            x = self.proj_in(x)
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, x.shape[1])
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states, num_view=num_view)
        if self.use_linear_projection:
            x = self.proj_out(x)
            x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)       # a channels-last view, no copy
            if not residual.is_contiguous(memory_format=torch.channels_last):
                x = x.contiguous()
        else:
            x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()
            x = self.proj_out(x)
        return x + residual


# --------------------------------------------------------------------------- conv blocks
class ResnetBlock2D(nn.Module):
    """resnet.py:200-376, default time-scale-shift, swish."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int], groups: int = 32,
                 eps: float = 1e-5, output_scale_factor: float = 1.0):
        super().__init__()
        self.norm1 = GroupNormAct(groups, in_channels, eps=eps, act=True)      # GroupNorm -> SiLU (:338-339)
        self.conv1 = BiasConv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = GroupNormAct(groups, out_channels, eps=eps, act=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = BiasConv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = BiasConv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self.output_scale_factor = output_scale_factor
        self.fused = False

    def _forward_fused(self, x, temb):
        """Same arithmetic with the element-wise passes folded away: conv1's bias and the projected time
        embedding enter GroupNorm-2 as a per-(sample, channel) pre-add, conv2's bias (+ shortcut bias) and the
        skip connection are one pass (resnet.py:338-376)."""
        from .. import ops
        if not _is_cl(x):
            x = x.contiguous(memory_format=torch.channels_last)
        h = F.conv2d(self.norm1(x), self.conv1.weight, None, padding=1)
        pre = self.conv1.bias[None].expand(x.shape[0], -1)
        if self.time_emb_proj is not None and temb is not None:
            pre = self.time_emb_proj(self.nonlinearity(temb)) + self.conv1.bias
        h = ops.group_norm_nhwc(h, self.norm2.num_groups, self.norm2.weight, self.norm2.bias, self.norm2.eps, True, pre_add=pre)
        h = F.conv2d(h, self.conv2.weight, None, padding=1)
        bias = self.conv2.bias
        if self.conv_shortcut is not None:
            x = F.conv2d(x, self.conv_shortcut.weight, None)
            bias = bias + self.conv_shortcut.bias
        out = ops.bias_residual_nhwc(h, bias=bias, residual=x)
        return out if self.output_scale_factor == 1.0 else out / self.output_scale_factor

    def forward(self, x, temb=None):
        if (self.fused and x.is_cuda and x.dim() == 4 and self.norm2._fusable(x)
                and _is_cl(self.conv1.weight) and self.conv1.weight.shape[2] == 3):
            return self._forward_fused(x, temb)
        h = self.conv1(self.norm1(x))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        out = x + h
        return out if self.output_scale_factor == 1.0 else out / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.padding = padding
        self.conv = BiasConv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


def subpixel_phase_weights(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, 3, 3] -> [4 C_out, C_in, 2, 2]: a 3x3 convolution applied to a nearest-2x-upsampled image, rewritten on
    the low-resolution image.  Output pixel (2i + a, 2j + b) reads upsampled rows 2i + a - 1 .. 2i + a + 1 = source rows
    {i - 1, i, i} for a = 0 and {i, i, i + 1} for a = 1 (same for columns), so phase (a, b) is a 2x2 convolution over source
    rows (i - 1 + a, i + a) whose taps are sums of the 3x3 taps that land on the same source pixel (zero padding of the upsampled
    image = zero padding of the source).  Channel block 2a + b holds phase (a, b).  Sums in fp32, one rounding to w.dtype."""
    rows = (((0,), (1, 2)), ((0, 1), (2,)))                     # rows[a][u] = 3x3 tap rows that read source row i - 1 + a + u
    wf = w.detach().to(torch.float64 if w.dtype == torch.float64 else torch.float32)
    out = wf.new_zeros((4,) + tuple(w.shape[:2]) + (2, 2))
    for a in range(2):
        for b in range(2):
            for u in range(2):
                for v in range(2):
                    acc = 0
                    for dy in rows[a][u]:
                        for dx in rows[b][v]:
                            acc = acc + wf[:, :, dy, dx]
                    out[2 * a + b, :, :, u, v] = acc
    return out.reshape(4 * w.shape[0], w.shape[1], 2, 2).to(w.dtype)


class Upsample2D(nn.Module):
    """upsampling.py Upsample2D (nearest, then a 3x3 convolution).  With `fused = True` on CUDA channels-last tensors the exact-2x
    case runs in sub-pixel form: one 2x2 cuDNN convolution on the low-resolution tensor with the four phase kernels of
    `subpixel_phase_weights` (2.25x fewer multiply-adds, no upsampled temporary) and rd_phase_shuffle_bias_nhwc interleaving the
    phases and adding the bias; other output sizes go through rd_upsample_nearest_nhwc + the 3x3 convolution."""

    subpixel = True          # class-wide switch (bench.py --subpixel 0 for the A/B)

    def __init__(self, channels: int):
        super().__init__()
        self.conv = BiasConv2d(channels, channels, 3, padding=1)
        self.fused = False
        self._phase_w = None
        self._phase_key = None

    def phase_weight(self) -> torch.Tensor:
        w = self.conv.weight
        key = (w.data_ptr(), w._version, w.dtype, w.device)
        if self._phase_key != key:
            self._phase_w = subpixel_phase_weights(w).contiguous(memory_format=torch.channels_last)
            self._phase_key = key
        return self._phase_w

    def forward(self, x, output_size=None):
        if self.fused and x.is_cuda and _is_cl(x) and x.shape[1] % (16 // x.element_size()) == 0:
            from .. import ops
            size = tuple(output_size) if output_size is not None else (2 * x.shape[2], 2 * x.shape[3])
            if self.subpixel and size == (2 * x.shape[2], 2 * x.shape[3]):
                t = F.conv2d(x, self.phase_weight(), None, 1, 1)
                if _is_cl(t):
                    return ops.phase_shuffle_bias_nhwc(t, self.conv.bias, x.shape[2], x.shape[3])
            return self.conv(ops.upsample_nearest_nhwc(x, size))
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=tuple(output_size), mode="nearest")
        return self.conv(x)
