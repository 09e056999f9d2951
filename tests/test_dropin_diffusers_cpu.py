"""Drop-in boundary of the attention processor on the REFERENCE's own diffusers tree (build container only).

`B200XFrameAttnProcessor` is installed with `unet.set_attn_processor` on a real `diffusers.UNet2DConditionModel`
(the fork vendored under /root/reference/diffusers, with the RollingDepth `num_view` edits) and driven through the real
`Attention.forward` (kwarg filtering by signature, attention_processor.py:483-500) and the real recursion of
`set_attn_processor` (unets/unet_2d_condition.py:721-753).  There is no GPU here, so the ONE call into the CUDA library
(`ops.xattn_fwd`) is replaced by the oracle's softmax(QK^T)V for the duration of the test: everything else -- the
`num_view` rearranges, the projections, the text cross-attention with a batch-1 context, the residual handling -- is
the product processor's host code, and the UNet output must equal the reference's own `AttnProcessor2_0`.

Skipped where /root/reference does not exist (the GPU box); the GPU-side equivalent on our own SD2 modules is
tests/test_gpu_parity.py::test_processor_* and tests/test_gpu_pipeline.py.
"""
import os
import sys

import pytest
import torch

import model_utils as mu

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(f"{REF}/diffusers/src"), reason="reference tree not present")


@pytest.fixture(scope="module")
def diffusers_ref():
    from oracle.gen_golden import import_reference
    import_reference()
    import diffusers
    assert diffusers.__file__.startswith(REF)
    return diffusers


@pytest.fixture()
def cpu_kernel_stub(monkeypatch):
    """Replace the C-ABI attention call by the oracle (test-only; the product has no CPU path)."""
    from oracle import xframe_attention as oxa
    from rollingdepth_b200 import _lib as L
    from rollingdepth_b200 import ops
    calls = []

    def fake_xattn(q, k, v, scale=None, out=None, flags=0):
        calls.append((tuple(q.shape), tuple(k.shape)))
        o = oxa.sdpa_ref(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float(), scale=scale)
        return o.transpose(1, 2).to(q.dtype)

    monkeypatch.setattr(ops, "xattn_fwd", fake_xattn)
    monkeypatch.setattr(L, "require_cuda", lambda *a, **k: None)
    return calls


def test_processor_on_real_diffusers_unet(diffusers_ref, cpu_kernel_stub):
    from diffusers import UNet2DConditionModel
    from diffusers.models.attention_processor import Attention, AttnProcessor2_0
    from rollingdepth_b200.attn_processor import B200XFrameAttnProcessor
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
    g = torch.Generator().manual_seed(4)
    n = 3
    x = torch.randn(2 * n, 8, 8, 12, generator=g)               # two snippets of three frames in one call
    t = torch.full((2 * n,), 999, dtype=torch.int64)
    ctx = torch.randn(1, 2, 32, generator=g)
    with torch.no_grad():
        unet.set_attn_processor(AttnProcessor2_0())
        # the reference processor needs the context expanded to the snippet batch
        ref = unet(x, t, encoder_hidden_states=ctx.expand(2, -1, -1), num_view=n).sample
        ref_pf = unet(x, t, encoder_hidden_states=ctx.expand(2 * n, -1, -1), num_view=None).sample
        proc = B200XFrameAttnProcessor()
        unet.set_attn_processor(proc)                           # the real recursion over all Attention modules
        attn_mods = [m for m in unet.modules() if isinstance(m, Attention)]
        assert len(attn_mods) == len(unet.attn_processors) > 0
        assert all(m.processor is proc for m in attn_mods) and all(p is proc for p in unet.attn_processors.values())
        out = unet(x, t, encoder_hidden_states=ctx, num_view=n).sample          # batch-1 context accepted
        out_b = unet(x, t, encoder_hidden_states=ctx.expand(2, -1, -1), num_view=n).sample
        out_pf = unet(x, t, encoder_hidden_states=ctx, num_view=None).sample
    # Attention.forward delivered `num_view` (a processor that does not name it is silently degraded to per-frame)
    sq = {c[0][1] for c in cpu_kernel_stub}
    assert n * 8 * 12 in sq, sq                                  # cross-frame: n * hw query tokens at level 0
    assert float((out - ref).abs().max()) < 2e-5
    assert float((out_b - ref).abs().max()) < 2e-5
    assert float((out_pf - ref_pf).abs().max()) < 2e-5
    assert float((ref - ref_pf).abs().max()) > 1e-4              # the cross-frame path matters on this input


def test_attention_forward_filters_unknown_kwargs_but_keeps_num_view(diffusers_ref, cpu_kernel_stub):
    """attention_processor.py:483-492: kwargs the processor's signature does not name are dropped with a warning;
    `num_view` must survive because `B200XFrameAttnProcessor.__call__` names it."""
    import inspect
    from diffusers.models.attention_processor import Attention
    from rollingdepth_b200.attn_processor import B200XFrameAttnProcessor
    assert "num_view" in inspect.signature(B200XFrameAttnProcessor.__call__).parameters
    attn = Attention(query_dim=64, heads=1, dim_head=64, processor=B200XFrameAttnProcessor())
    x = torch.randn(6, 5, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y = attn(x, num_view=3, some_unknown_kwarg=1)
        y_pf = attn(x)
    assert y.shape == x.shape and cpu_kernel_stub[0][0][:2] == (2, 15) and cpu_kernel_stub[1][0][:2] == (6, 5)
    assert float((y - y_pf).abs().max()) > 1e-4


def test_processor_on_real_diffusers_vae_is_left_alone(diffusers_ref):
    """The product pipeline installs the processor on the UNet only (leave the VAE on its default)."""
    from diffusers import AutoencoderKL
    from rollingdepth_b200.attn_processor import B200XFrameAttnProcessor
    vae = AutoencoderKL(**mu.TINY_VAE)
    assert not any(isinstance(p, B200XFrameAttnProcessor) for p in vae.attn_processors.values())
