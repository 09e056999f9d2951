"""GPU parity tests (run by the driver with `-m gpu` on a B200).

Every check goes through the C ABI (rollingdepth_b200._lib -> librd_b200.so) and compares with
(a) fixtures produced by the reference's own code (tests/golden, see oracle/gen_golden.py) and
(b) the CPU oracle on the same seeded inputs.  Tolerances are the north star's:
attention <= 2e-2 max-abs (fp16/bf16) and <= 1e-4 (fp32); co-alignment scale/shift <= 1e-4 relative,
checked teacher-forced per iteration; integer/byte-exact pieces (min, crop, merge in fp16, DDIM, renorm)
are compared bit for bit.
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import coalign as oc
from oracle import schedule as osch
from oracle import xframe_attention as oxa


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch.device("cuda:0")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# =============================================================================== library
def test_library_loads_on_b200():
    from rollingdepth_b200 import _lib as L
    _dev()
    assert L.lib().rd_version() == 1
    assert L.lib().rd_device_arch() == 100, "kernels are built for sm_100a only"


# =============================================================================== attention kernel
def _rand_qkv(B, Sq, Sk, H, dtype, dev, seed=0, packed=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    D = 64
    if packed:  # the reference layout: [B, S, H*D] viewed as [B, S, H, D]
        q = torch.randn(B, Sq, H * D, generator=g).to(dtype).to(dev).view(B, Sq, H, D)
        k = torch.randn(B, Sk, H * D, generator=g).to(dtype).to(dev).view(B, Sk, H, D)
        v = torch.randn(B, Sk, H * D, generator=g).to(dtype).to(dev).view(B, Sk, H, D)
    else:       # head-major storage, exercised through strides
        q = torch.randn(B, H, Sq, D, generator=g).to(dtype).to(dev).transpose(1, 2)
        k = torch.randn(B, H, Sk, D, generator=g).to(dtype).to(dev).transpose(1, 2)
        v = torch.randn(B, H, Sk, D, generator=g).to(dtype).to(dev).transpose(1, 2)
    return q, k, v


def _oracle_attn(q, k, v, scale=None):
    # oracle evaluated in fp32 on the same device (the score matrix is chunked)
    o = oxa.sdpa_ref(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float(), scale=scale)
    return o.transpose(1, 2)


XATTN_SHAPES = [
    # B, Sq, Sk, H          what it covers
    (1, 128, 128, 1),      # exactly one tile
    (1, 252, 252, 20),     # fast preset mid block (S = 3*84), ragged tail
    (1, 1008, 1008, 20),   # fast level 2
    (1, 3888, 3888, 10),   # fast level 1
    (2, 300, 300, 5),      # batch > 1 (snippet batching)
    (1, 1000, 2, 5),       # text cross-attention: 2 key tokens (BOS+EOS)
    (1, 77, 333, 3),       # Sq != Sk, both ragged
    (1, 1, 1, 1),          # degenerate
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", XATTN_SHAPES)
@pytest.mark.parametrize("flags", [0, 2, 4 << 8, 2 << 4, 3 << 4, (3 << 4) | (4 << 8), (3 << 4) | (8 << 8)])
def test_xattn_tcgen05_matches_oracle(shape, dtype, flags):
    from rollingdepth_b200 import ops
    dev = _dev()
    B, Sq, Sk, H = shape
    q, k, v = _rand_qkv(B, Sq, Sk, H, dtype, dev, seed=Sq + Sk)
    out = ops.xattn_fwd(q, k, v, flags=flags)
    torch.cuda.synchronize()
    ref = _oracle_attn(q, k, v)
    err = float((out.float() - ref).abs().max())
    assert out.shape == (B, Sq, H, 64) and out.dtype == dtype
    assert math.isfinite(err) and err <= 2e-2, f"max-abs {err}"
    # mean error is far below the bound when the kernel is right; catches systematic layout slips
    assert float((out.float() - ref).abs().mean()) <= 2e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_xattn_full_size_level0_fast_preset(dtype):
    """S = 3 * 54 * 96 = 15552 tokens, 5 heads: the largest layer of the `fast` preset."""
    from rollingdepth_b200 import ops
    dev = _dev()
    q, k, v = _rand_qkv(1, 15552, 15552, 5, dtype, dev, seed=5)
    out = ops.xattn_fwd(q, k, v)
    ref = _oracle_attn(q, k, v)
    err = float((out.float() - ref).abs().max())
    assert err <= 2e-2, f"max-abs {err}"


def _sampled_rows_ref(q, k, v, b, h, rows):
    """fp32 softmax(q k^T / sqrt(D)) v for a few query rows of one (batch, head) -- exact reference for shapes whose
    full score matrix would not be worth computing."""
    qs = q[b, rows, h].float()                                   # [r, D]
    s = (qs @ k[b, :, h].float().T) * (q.shape[-1] ** -0.5)      # [r, Sk]
    return torch.softmax(s, dim=-1) @ v[b, :, h].float()


@pytest.mark.parametrize("B,S,H,what", [
    (16, 15552, 5, "fast preset level 0 at the snippet batch the pipeline runs (auto-selects the three-tile kernel)"),
    (4, 27648, 5, "full preset level 0, many-wave grid (three-tile kernel)"),
    (16, 3888, 10, "fast preset level 1 at B = 16"),
    (1, 131072, 5, "largest shape of the config-5 sweep: n = 8 frames of 1024x1024"),
    (1, 49152, 5, "n = 3 frames of 1024x1024"),
])
def test_xattn_production_shapes_default_flags(B, S, H, what):
    """flags = 0 (the kernel selection the product path uses) at the shapes the bench runs, checked on sampled
    (batch, head, row) triples against an fp32 softmax over all keys."""
    from rollingdepth_b200 import ops
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(S + B)
    q, k, v = [torch.randn(B, S, H * 64, device=dev, dtype=torch.float16, generator=g).view(B, S, H, 64) for _ in range(3)]
    out = ops.xattn_fwd(q, k, v, flags=0)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    worst = 0.0
    rows = torch.tensor([0, 1, 127, 128, 383, 384, 385, S // 2, S - 385, S - 129, S - 2, S - 1], device=dev)
    for b in sorted({0, B // 2, B - 1}):
        for h in sorted({0, H // 2, H - 1}):
            ref = _sampled_rows_ref(q, k, v, b, h, rows)
            worst = max(worst, float((out[b, rows, h].float() - ref).abs().max()))
    assert worst <= 2e-2, f"{what}: max-abs {worst}"
    # every query tile was written: no row of the (poisoned) output may be left untouched
    out2 = torch.full_like(out, float("nan"))
    ops.xattn_fwd(q, k, v, out=out2, flags=0)
    assert torch.isfinite(out2.float()).all()
    assert torch.equal(out, out2), "the kernel must be deterministic run to run"


@pytest.mark.parametrize("f32_path", [1, 2], ids=["ffma", "3xtf32"])
def test_xattn_fp32_processor_shape_level0(f32_path):
    """fp32 path (`paper` preset / cfg1) at S = 15552, 5 heads: <= 1e-4 max-abs against a float64 softmax on sampled rows."""
    from rollingdepth_b200 import ops, _lib as L
    dev = _dev()
    S, H = 15552, 5
    g = torch.Generator(device=dev).manual_seed(7)
    q, k, v = [torch.randn(1, S, H * 64, device=dev, dtype=torch.float32, generator=g).view(1, S, H, 64) for _ in range(3)]
    out = ops.xattn_fwd(q, k, v, flags=L.XATTN_F32(f32_path))
    assert torch.isfinite(out).all()
    rows = torch.tensor([0, 77, 5000, S - 1], device=dev)
    for h in (0, 4):
        s = (q[0, rows, h].double() @ k[0, :, h].double().T) * 0.125
        ref = torch.softmax(s, dim=-1) @ v[0, :, h].double()
        assert float((out[0, rows, h].double() - ref).abs().max()) <= 1e-4


def test_xattn_head_major_strides_and_scale():
    from rollingdepth_b200 import ops
    dev = _dev()
    q, k, v = _rand_qkv(2, 200, 456, 4, torch.float16, dev, seed=9, packed=False)
    out = ops.xattn_fwd(q, k, v, scale=0.2)
    ref = _oracle_attn(q, k, v, scale=0.2)
    assert float((out.float() - ref).abs().max()) <= 2e-2


@pytest.mark.parametrize("flags", [0, 3 << 4, (3 << 4) | (4 << 8)])
def test_xattn_large_logits_trigger_rescale(flags):
    """Scores that keep growing along the key axis force the lazy O-rescale path on every tile (both kernels)."""
    from rollingdepth_b200 import ops
    dev = _dev()
    B, S, H, D = 1, 1024, 2, 64
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, S, H, D, generator=g)
    k = torch.randn(B, S, H, D, generator=g) * torch.linspace(0.5, 6.0, S)[None, :, None, None]
    v = torch.randn(B, S, H, D, generator=g)
    q, k, v = [t.half().to(dev) for t in (q, k, v)]
    out = ops.xattn_fwd(q, k, v, flags=flags)
    ref = _oracle_attn(q, k, v)
    assert float((out.float() - ref).abs().max()) <= 2e-2


def test_xattn_rejects_bad_arguments():
    from rollingdepth_b200 import ops, _lib as L
    dev = _dev()
    q = torch.randn(1, 16, 2, 32, device=dev, dtype=torch.float16)
    with pytest.raises(L.RdError):
        ops.xattn_fwd(q, q, q)            # head_dim 32 is not instantiated
    with pytest.raises(L.RdError):
        ops.xattn_fwd(q.cpu(), q.cpu(), q.cpu())   # no CPU fallback


@pytest.mark.parametrize("f32_path", [0, 1, 2], ids=["default", "ffma", "3xtf32"])
@pytest.mark.parametrize("shape", [(1, 252, 252, 20), (1, 1024, 1024, 5), (2, 130, 67, 3), (1, 500, 2, 5), (1, 1, 1, 1), (2, 3072, 3072, 5)])
def test_xattn_fp32_matches_oracle(shape, f32_path):
    """fp32 inputs: the scalar FFMA kernel and the error-compensated 3xTF32 tcgen05 kernel, both within the north star's
    1e-4 max-abs of a float64 softmax."""
    from rollingdepth_b200 import ops, _lib as L
    dev = _dev()
    B, Sq, Sk, H = shape
    q, k, v = _rand_qkv(B, Sq, Sk, H, torch.float32, dev, seed=3)
    out = ops.xattn_fwd(q, k, v, flags=L.XATTN_F32(f32_path))
    ref = oxa.sdpa_ref(q.transpose(1, 2).double(), k.transpose(1, 2).double(), v.transpose(1, 2).double(),
                       acc_dtype=torch.float64).transpose(1, 2)
    assert float((out.double() - ref).abs().max()) <= 1e-4


# =============================================================================== processor vs reference fixtures
class _Attn(torch.nn.Module):
    """Minimal stand-in exposing what a diffusers `Attention` exposes to its processor
    (attention_processor.py:100-268): to_q/to_k/to_v (bias-free), to_out[0..1], heads."""

    def __init__(self, g, tag, dev, dtype):
        super().__init__()
        T = lambda k: torch.from_numpy(g[f"{tag}_{k}"]).to(dev).to(dtype)
        C, heads = int(g[f"{tag}_meta"][0]), int(g[f"{tag}_meta"][1])
        cin_kv = T("wk").shape[1]
        self.heads = heads
        self.to_q = torch.nn.Linear(C, heads * 64, bias=False)
        self.to_k = torch.nn.Linear(cin_kv, heads * 64, bias=False)
        self.to_v = torch.nn.Linear(cin_kv, heads * 64, bias=False)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(heads * 64, C), torch.nn.Dropout(0.0)])
        self.to(dev).to(dtype)
        with torch.no_grad():
            self.to_q.weight.copy_(T("wq")); self.to_k.weight.copy_(T("wk")); self.to_v.weight.copy_(T("wv"))
            self.to_out[0].weight.copy_(T("wo")); self.to_out[0].bias.copy_(T("bo"))
        self.spatial_norm = self.group_norm = self.norm_q = self.norm_k = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0


@pytest.mark.parametrize("tag", ["self", "self_odd", "text", "perframe"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.float16, 2e-2)])
def test_processor_matches_reference_fixture(golden_dir, tag, dtype, tol):
    from rollingdepth_b200.attn_processor import B200XFrameAttnProcessor
    dev = _dev()
    g = _load(golden_dir, "xattn.npz")
    C, heads, n, hw, cross, nv = [int(v) for v in g[f"{tag}_meta"]]
    attn = _Attn(g, tag, dev, dtype)
    x = torch.from_numpy(g[f"{tag}_x"]).to(dev).to(dtype)
    ctx = torch.from_numpy(g[f"{tag}_ctx"]).to(dev).to(dtype) if cross else None
    with torch.no_grad():
        y = B200XFrameAttnProcessor()(attn, x, encoder_hidden_states=ctx, num_view=None if nv < 0 else nv)
    ref = torch.from_numpy(g[f"{tag}_y"]).to(dev)
    assert y.shape == ref.shape and y.dtype == dtype
    assert float((y.float() - ref).abs().max()) <= tol


# =============================================================================== DDIM
@pytest.mark.parametrize("tag,ptype", [("vtrail", "v_prediction"), ("epslead", "epsilon")])
def test_ddim_step_vs_reference(golden_dir, tag, ptype):
    """Bit-exact against the oracle evaluated with the same torch ops on CUDA tensors (the roundings
    the reference's GPU path performs: fp32 schedule scalars times fp16 tensors, one rounding per op),
    and within one fp16 ulp of the reference's CPU fixtures (torch's CPU path rounds the 0-dim
    schedule scalar to fp16 first, its CUDA path does not)."""
    from rollingdepth_b200.scheduler import DDIMSchedule
    dev = _dev()
    g = _load(golden_dir, "ddim.npz")
    kw = dict(prediction_type=ptype, timestep_spacing="trailing" if tag == "vtrail" else "leading", steps_offset=1)
    for n in (1, 10, 20):
        sch = DDIMSchedule(**kw)
        orc = osch.DDIMRef(**kw)
        ts = sch.set_timesteps(n)
        orc.set_timesteps(n)
        assert ts == [int(t) for t in g[f"{tag}_ts{n}"]]
        for dn, dt in (("f32", torch.float32), ("f16", torch.float16)):
            for pos in sorted({0, len(ts) - 1}):
                x = torch.from_numpy(g[f"{tag}_n{n}_{dn}_p{pos}_x"]).to(dev)
                v = torch.from_numpy(g[f"{tag}_n{n}_{dn}_p{pos}_v"]).to(dev)
                y = sch.step(v, ts[pos], x)
                assert torch.equal(y, orc.step(v, ts[pos], x)), (tag, n, dn, pos)
                ref = torch.from_numpy(g[f"{tag}_n{n}_{dn}_p{pos}_y"]).to(dev)
                # CPU fixture: a few fp16 ulps (the CPU path rounds the schedule scalars to fp16 first)
                rel = 2e-6 if dt == torch.float32 else 8 * 2.0 ** -11
                assert float(((y.float() - ref.float()).abs() / ref.float().abs().clamp_min(1.0)).max()) <= rel
    sch = DDIMSchedule(**kw)
    sch.set_timesteps(20)
    for dn in ("f32", "f16"):
        z = sch.add_noise(torch.from_numpy(g[f"{tag}_addnoise_{dn}_x"]).to(dev),
                          torch.from_numpy(g[f"{tag}_addnoise_{dn}_n"]).to(dev), int(g[f"{tag}_addnoise_t"]))
        ref = torch.from_numpy(g[f"{tag}_addnoise_{dn}_y"]).to(dev)
        assert torch.equal(z, ref)


# =============================================================================== co-alignment
def _case(g, tag):
    meta = [int(v) for v in g[f"{tag}_meta"]]
    N, H, W, w, iters, G = meta[:6]
    dil = meta[6:6 + G]
    snips = [torch.from_numpy(g[f"{tag}_snip{i}"]) for i in range(G)]
    return N, H, W, w, iters, G, dil, snips


@pytest.fixture(params=[-1, 0, 1, 2, 3, 4, 5, 6], ids=lambda c: f"cfg{c}")
def coalign_cfg(request):
    """Every organisation of the optimisation pass (include/rd_b200.h rd_coalign_set_config): -1 automatic,
    0 / 1 pixel-packed kernel, 2 / 3 / 4 contribution-packed kernel, 5 / 6 frame-persistent contribution-packed kernel."""
    from rollingdepth_b200 import _lib as L
    assert L.lib().rd_coalign_set_config(request.param) >= 1
    yield request.param
    L.lib().rd_coalign_set_config(-1)


@pytest.mark.parametrize("tag", ["a", "b", "h", "w2"])
def test_min_and_prep_bit_exact(golden_dir, tag):
    from rollingdepth_b200.depth_aligner import DepthAligner
    dev = _dev()
    g = _load(golden_dir, "coalign.npz")
    N, H, W, w, iters, G, dil, snips = _case(g, tag)
    al = DepthAligner(device=dev)
    pb, dsn, dev_min, x_small = al.prepare(snips, dil)
    mn, _, small = oc.prep(snips)
    assert float(dev_min.cpu()) == float(mn)
    ours = x_small[:, :, :pb.P].cpu()
    ref = torch.cat(small, 0)
    assert ours.dtype == ref.dtype and torch.equal(ours, ref)
    assert float(x_small[:, :, pb.P:].abs().max().cpu() if pb.P_pad > pb.P else 0.0) == 0.0


@pytest.mark.parametrize("tag", ["a", "b", "h", "w2"])
def test_coalign_teacher_forced_one_iteration(golden_dir, tag, coalign_cfg):
    """Load the reference's (s, t, m, v, step) at iteration k, run ONE of our iterations, compare the
    new parameters with the reference's own autograd + torch.optim.Adam result and the loss entry."""
    from rollingdepth_b200.depth_aligner import DepthAligner
    dev = _dev()
    g = _load(golden_dir, "coalign.npz")
    N, H, W, w, iters, G, dil, snips = _case(g, tag)
    al = DepthAligner(device=dev)
    pb, dsn, dev_min, x_small = al.prepare(snips, dil)
    for it in sorted({0, 1, 5, iters - 1}):
        cat = lambda key, idx: torch.cat([torch.from_numpy(g[f"{tag}_it{it}_{key}{i}"]) for i in idx]).to(dev)
        s, t = cat("pre", range(G)), cat("pre", range(G, 2 * G))
        if it > 0:
            ms, mt = cat("m", range(G)), cat("m", range(G, 2 * G))
            vs, vt = cat("v", range(G)), cat("v", range(G, 2 * G))
        else:
            ms, mt, vs, vt = [torch.zeros_like(s) for _ in range(4)]
        state = [s.clone(), t.clone(), ms.clone(), mt.clone(), vs.clone(), vt.clone()]
        hist = al.optimize_state(pb, x_small, state, step0=it, n_iter=1)
        ref_s, ref_t = cat("post", range(G)), cat("post", range(G, 2 * G))
        ref_hist = g[f"{tag}_hist"][it]
        h = hist.cpu().numpy()[0]
        assert abs(h[0] - ref_hist[0]) <= 1e-5 * max(1.0, abs(ref_hist[0])), (h, ref_hist)
        assert abs(h[1] - ref_hist[1]) <= 1e-5 and abs(h[2] - ref_hist[2]) <= 1e-5
        # scale/shift within 1e-4 relative (north star); the step itself is ~lr = 1e-3
        assert float(((state[0] - ref_s).abs() / ref_s.abs().clamp_min(1e-3)).max()) <= 1e-4
        assert float((state[0] - ref_s).abs().max()) <= 5e-6
        assert float((state[1] - ref_t).abs().max()) <= 5e-6


@pytest.mark.parametrize("tag", ["a", "b", "w2"])
def test_coalign_free_run_matches_reference_history(golden_dir, tag, coalign_cfg):
    from rollingdepth_b200.depth_aligner import DepthAligner
    dev = _dev()
    g = _load(golden_dir, "coalign.npz")
    N, H, W, w, iters, G, dil, snips = _case(g, tag)
    al = DepthAligner(device=dev, num_iterations=iters)
    merged, scales, trans, hist = al.run(snips, dil)
    ref_hist = g[f"{tag}_hist"]
    assert len(hist) == iters
    np.testing.assert_allclose(np.array(hist)[:8], ref_hist[:8], rtol=2e-5, atol=2e-6)
    # the sign-gradient L1 + Adam dynamics are chaotic beyond ~10 iterations (hard part 2):
    # require agreement of the final loss, not of individual parameters
    assert abs(hist[-1][0] - ref_hist[-1][0]) <= 2e-3 * abs(ref_hist[-1][0])
    assert merged.shape == (N, 1, H, W)
    for i in range(G):
        assert scales[i].shape == (g[f"{tag}_scale{i}"].shape[0], 1, 1) and scales[i].dtype == torch.float32


@pytest.mark.parametrize("tag", ["a", "b", "h", "w2"])
def test_merge_matches_reference(golden_dir, tag):
    from rollingdepth_b200.depth_aligner import DepthAligner
    dev = _dev()
    g = _load(golden_dir, "coalign.npz")
    N, H, W, w, iters, G, dil, snips = _case(g, tag)
    al = DepthAligner(device=dev)
    pb, dsn, dev_min, _ = al.prepare(snips, dil)
    s = torch.cat([torch.from_numpy(g[f"{tag}_scale{i}"]) for i in range(G)]).to(dev)
    t = torch.cat([torch.from_numpy(g[f"{tag}_trans{i}"]) for i in range(G)]).to(dev)
    merged = al.merge(pb, dsn, dev_min, s, t).cpu()
    ref = torch.from_numpy(g[f"{tag}_merged"])
    assert merged.dtype == ref.dtype and merged.shape == ref.shape
    if ref.dtype == torch.float16:
        assert torch.equal(merged, ref)                     # same roundings => bit-equal
    else:
        assert float((merged - ref).abs().max()) <= 2e-7    # fp32: summation order of <= 9 terms
    mm = al.merged_minmax.cpu()
    assert float(mm[0]) == float(merged.float().min()) and float(mm[1]) == float(merged.float().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_renorm_bit_exact(dtype):
    from rollingdepth_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(2)
    d = (torch.rand(7, 1, 33, 40, generator=g) * 3 + 0.25).to(dtype)
    ref = oc.renorm(d.clone())
    dd = d.to(dev)
    mm = torch.stack([dd.float().min(), dd.float().max()])
    out = ops.renorm_(dd.clone(), mm).cpu()
    assert torch.equal(out, ref)


def test_coalign_fast_preset_size_vs_oracle(coalign_cfg):
    """N = 150 frames at 432x768 fp16, dilations [1, 25] (the `fast` preset's aligner problem):
    three iterations against the CPU oracle + invariants at full size."""
    from rollingdepth_b200.depth_aligner import DepthAligner
    from oracle.gen_golden import synth_snippets
    dev = _dev()
    N, H, W, w, dil = 150, 432, 768, 3, [1, 25]
    snips = synth_snippets(N, H, W, w, dil, seed=1, dtype=torch.float16)
    al = DepthAligner(device=dev, num_iterations=3)
    pb, dsn, dev_min, x_small = al.prepare(snips, dil)
    assert (pb.Hs, pb.Ws, pb.P) == (43, 77, 3311)
    state = al.new_state(pb)
    hist = al.optimize_state(pb, x_small, state, 0, 3).cpu().numpy()
    _, shifted, small = oc.prep(snips)
    s_ls, t_ls, ohist, _ = oc.optimize(small, dil, N, num_iterations=3)
    np.testing.assert_allclose(hist, np.array(ohist), rtol=2e-5, atol=2e-6)
    # sign-gradient + Adam: a single pixel whose residual sign differs moves a step by a fraction of lr = 1e-3;
    # the north-star bound is 1e-4 relative
    assert float(((state[0].cpu() - torch.cat(s_ls)).abs() / torch.cat(s_ls).abs()).max()) <= 1e-4
    assert float((state[1].cpu() - torch.cat(t_ls)).abs().max()) <= 1e-4
    # merge invariant: with s = 1, t = 0 every frame is the plain mean of its (shifted) predictions
    merged = al.merge(pb, dsn, dev_min, torch.ones_like(state[0]), torch.zeros_like(state[1])).cpu()
    ref = oc.merge(shifted, [torch.ones(n) for n in (148, 100)], [torch.zeros(n) for n in (148, 100)], dil, N)
    assert torch.equal(merged, ref)


def test_coalign_multi_chunk_frames_vs_oracle(coalign_cfg):
    """1024x576 fp32 with three dilations (R = 9 contributions, P = 5916 subsampled pixels): every frame is split into
    several chunks, so the cross-chunk hand-over (ticket, last chunk finalises the frame) and the boundary frames with
    absent contributions are exercised; three iterations against the CPU oracle, and two runs must agree bit for bit."""
    from rollingdepth_b200.depth_aligner import DepthAligner
    from oracle.gen_golden import synth_snippets
    dev = _dev()
    N, H, W, w, dil = 24, 576, 1024, 3, [1, 4, 9]
    snips = synth_snippets(N, H, W, w, dil, seed=5)
    al = DepthAligner(device=dev, num_iterations=3)
    pb, dsn, dev_min, x_small = al.prepare(snips, dil)
    assert pb.P == 58 * 102
    state = al.new_state(pb)
    hist = al.optimize_state(pb, x_small, state, 0, 3).cpu().numpy()
    state2 = al.new_state(pb)
    hist2 = al.optimize_state(pb, x_small, state2, 0, 3).cpu().numpy()
    assert all(torch.equal(a, b) for a, b in zip(state, state2)) and np.array_equal(hist, hist2)
    _, _, small = oc.prep(snips)
    s_ls, t_ls, ohist, _ = oc.optimize(small, dil, N, num_iterations=3)
    np.testing.assert_allclose(hist, np.array(ohist), rtol=2e-5, atol=2e-6)
    assert float(((state[0].cpu() - torch.cat(s_ls)).abs() / torch.cat(s_ls).abs()).max()) <= 1e-4
    assert float((state[1].cpu() - torch.cat(t_ls)).abs().max()) <= 1e-4


def test_coalign_recovers_known_affine():
    """Property at moderate size: snippets that are exact affine images of one base video are
    re-aligned -- the loss falls by > 5x and the merged video correlates with the base."""
    from rollingdepth_b200.depth_aligner import DepthAligner
    from oracle.gen_golden import synth_snippets
    dev = _dev()
    N, H, W, w, dil = 40, 104, 104, 3, [1, 5]
    snips = synth_snippets(N, H, W, w, dil, seed=3, noise=0.0)
    al = DepthAligner(device=dev, num_iterations=600)
    merged, scales, trans, hist = al.run(snips, dil)
    assert hist[-1][0] < hist[0][0] / 5
    assert all(math.isfinite(h[0]) for h in hist)


# =============================================================================== fused GroupNorm (+SiLU), NHWC
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(3, 128, 37, 53), (2, 320, 27, 48), (1, 1920, 14, 24), (5, 512, 54, 96), (2, 2560, 7, 12),
                                   (2, 160, 9, 11),        # 500 threads per CTA: a partial last warp
                                   (2, 64, 1, 1),          # a single pixel
                                   (1, 128, 432, 768),     # one sample: 128 chunks
                                   (40, 128, 40, 56)])     # many samples: 7 chunks each = one wave
@pytest.mark.parametrize("act", [False, True])
@pytest.mark.parametrize("pre", [False, True])
def test_groupnorm_nhwc_matches_torch_fp32(shape, dtype, act, pre):
    """Floating-point kernel: compared with a plain PyTorch fp32 reference of the same op."""
    from rollingdepth_b200 import ops
    dev = _dev()
    N, C, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(N, C, H, W, generator=g) * 1.5 + torch.randn(1, C, 1, 1, generator=g) * 2).to(dtype).to(dev)
    x = x.contiguous(memory_format=torch.channels_last)
    wt = (1 + 0.2 * torch.randn(C, generator=g)).to(dtype).to(dev)
    bs = (0.3 * torch.randn(C, generator=g)).to(dtype).to(dev)
    pv = (torch.randn(N, C, generator=g) * 0.5).to(dtype).to(dev) if pre else None
    y = ops.group_norm_nhwc(x, 32, wt, bs, 1e-6, act, pre_add=pv)
    xin = x.float() if not pre else (x.float() + pv.float()[:, :, None, None]).to(dtype).float()
    ref = torch.nn.functional.group_norm(xin, 32, wt.float(), bs.float(), 1e-6)
    if act:
        ref = torch.nn.functional.silu(ref.to(dtype).float())
    assert y.shape == x.shape and y.is_contiguous(memory_format=torch.channels_last)
    tol = {torch.float16: 2e-3, torch.bfloat16: 1.6e-2, torch.float32: 2e-5}[dtype]
    err = ((y.float() - ref).abs() / ref.abs().clamp_min(1.0)).max()
    assert float(err) <= tol, float(err)
    # [N, L, C] entry (attention processors)
    x3 = x.permute(0, 2, 3, 1).reshape(N, H * W, C)
    y3 = ops.group_norm_nhwc(x3, 32, wt, bs, 1e-6, act, pre_add=pv)
    assert torch.equal(y3.reshape(N, H, W, C).permute(0, 3, 1, 2), y)


def test_fused_modules_match_plain_torch_modules():
    """The SD2 host modules with NHWC weights + fused norms against the same modules in plain torch."""
    import model_utils as mu
    from rollingdepth_b200.sd2.unet import UNet2DConditionModel
    from rollingdepth_b200.sd2.vae import AutoencoderKL
    dev = _dev()
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)

    def fuse(m):
        m.to(memory_format=torch.channels_last)
        for sub in m.modules():
            if hasattr(sub, "fused"):
                sub.fused = True

    for dtype, tol in ((torch.float32, 2e-4), (torch.float16, 2e-2)):
        vae = vae.to(dev, dtype)
        unet = unet.to(dev, dtype)
        for sub in list(vae.modules()) + list(unet.modules()):
            if hasattr(sub, "fused"):
                sub.fused = False
        z = torch.randn(3, 4, 21, 24, device=dev, dtype=dtype)
        x = torch.randn(3, 8, 9, 14, device=dev, dtype=dtype)
        ctx = torch.randn(1, 2, 32, device=dev, dtype=dtype)
        img = torch.rand(2, 3, 24, 40, device=dev, dtype=dtype) * 2 - 1
        t = torch.full((3,), 500, device=dev)
        with torch.no_grad():
            torch.backends.cudnn.allow_tf32 = False
            ref_d = vae.decoder(vae.post_quant_conv(z))
            ref_e = vae.quant_conv(vae.encoder(img))
            ref_u = unet(x, t, encoder_hidden_states=ctx, num_view=3).sample
            fuse(vae); fuse(unet)
            out_d = vae.decoder(vae.post_quant_conv(z))
            out_e = vae.quant_conv(vae.encoder(img))
            out_u = unet(x, t, encoder_hidden_states=ctx, num_view=3).sample
        for o, r in ((out_d, ref_d), (out_e, ref_e), (out_u, ref_u)):
            assert float((o.float() - r.float()).abs().max()) <= tol * max(1.0, float(r.float().abs().max()))


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_bias_residual_and_upsample_nhwc(dtype):
    from rollingdepth_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 320, 13, 17, generator=g).to(dtype).to(dev).contiguous(memory_format=torch.channels_last)
    r = torch.randn(3, 320, 13, 17, generator=g).to(dtype).to(dev).contiguous(memory_format=torch.channels_last)
    b = torch.randn(320, generator=g).to(dtype).to(dev)
    v = torch.randn(3, 320, generator=g).to(dtype).to(dev)
    y = ops.bias_residual_nhwc(x, bias=b, vec=v, residual=r, inplace=False)
    ref = (x.float() + b.float().view(1, -1, 1, 1) + v.float()[:, :, None, None] + r.float()).to(dtype)
    assert float((y.float() - ref.float()).abs().max()) <= (4e-3 if dtype == torch.float16 else 1e-6)
    for size in ((26, 34), (27, 33), (14, 24)):
        up = ops.upsample_nearest_nhwc(x, size)
        assert torch.equal(up, torch.nn.functional.interpolate(x, size=size, mode="nearest"))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_geglu_matches_torch(dtype):
    """rd_geglu vs the reference's two-op form (activations.py GEGLU.forward) evaluated by torch on the same device."""
    from rollingdepth_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    for shape in [(3, 77, 2 * 1280), (2, 5, 2 * 64), (1, 1000, 2 * 320)]:
        proj = (torch.randn(shape, generator=g) * 2.0).to(dev, dtype)
        h, gate = proj.chunk(2, dim=-1)
        ref = h * torch.nn.functional.gelu(gate)
        out = ops.geglu(proj)
        assert out.shape == ref.shape and out.dtype == ref.dtype
        if dtype == torch.float32:
            torch.testing.assert_close(out, ref, rtol=2e-6, atol=2e-6)
        else:
            # same operation order and roundings as torch: at most one ulp from erff differences
            ulp = 2.0 ** (-10 if dtype == torch.float16 else -7)
            assert float(((out.float() - ref.float()).abs() / ref.float().abs().clamp_min(1e-3)).max()) <= 2 * ulp


# =============================================================================== VAE mid-block attention (head_dim 512)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("rows,cols", [(7, 5184), (3, 9216), (33, 36), (2, 10240), (5, 1300)])
def test_softmax_rows_matches_torch(rows, cols, dtype):
    """rd_softmax_rows = torch.softmax(scale * logits, -1).to(dtype) in one pass (fp32 math, ex2.approx)."""
    from rollingdepth_b200 import ops
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(rows * cols)
    x = torch.randn(rows, cols, device=dev, generator=g) * 30.0
    x[0, :4] = 200.0                                              # large logits: the row maximum must be subtracted
    scale = 512 ** -0.5
    y = ops.softmax_rows(x, scale, dtype)
    ref = torch.softmax(x.double() * scale, dim=-1)
    assert y.dtype == dtype and y.shape == x.shape
    tol = {torch.float32: 2e-6, torch.float16: 6e-4, torch.bfloat16: 5e-3}[dtype]
    assert float((y.double() - ref).abs().max()) <= tol
    assert float((y.double().sum(-1) - 1).abs().max()) <= 50 * tol


def test_softmax_rows_rejects_unsupported_rows():
    from rollingdepth_b200 import ops, _lib as L
    dev = _dev()
    with pytest.raises(L.RdError):
        ops.softmax_rows(torch.zeros(2, 10244, device=dev), 1.0, torch.float16)      # too many columns
    with pytest.raises(L.RdError):
        ops.softmax_rows(torch.zeros(2, 1302, device=dev), 1.0, torch.float16)       # not a multiple of 4


@pytest.mark.parametrize("S", [336, 5184])
def test_vae_mid_block_attention_matches_sdpa(S):
    """The VAE mid-block attention (one head, head_dim 512) as run by the sd2 modules -- cuBLAS batched GEMM with fp32
    logits, rd_softmax_rows, cuBLAS GEMM -- against F.scaled_dot_product_attention (the reference's call) on the same
    q, k, v."""
    from rollingdepth_b200.sd2 import layers as Ly
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(S)
    q, k, v = [torch.randn(2, 1, S, 512, device=dev, dtype=torch.float16, generator=g) for _ in range(3)]
    assert Ly._use_gemm_attention(q)
    o = Ly._wide_head_attention(q, k, v)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    assert o is not None and o.shape == q.shape
    assert float((o.float() - ref).abs().max()) <= 5e-3


# =============================================================================== LayerNorm of the transformer blocks
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2), (torch.float32, 2e-6)])
@pytest.mark.parametrize("M,C", [(1000, 320), (777, 640), (300, 1280), (5, 8), (33, 2048)])
def test_layernorm_matches_torch_fp32(M, C, dtype, tol):
    """Floating-point kernel: compared with F.layer_norm evaluated in fp32 on the same inputs; and with torch's own kernel in
    the activation dtype (<= 1 ulp of the output)."""
    from rollingdepth_b200 import ops
    dev = _dev()
    if dtype == torch.float32 and C > 1024:
        pytest.skip("fp32 rows hold at most 1024 channels")
    g = torch.Generator(device=dev).manual_seed(M + C)
    x = (torch.randn(M, C, device=dev, generator=g) * 2 + torch.randn(M, 1, device=dev, generator=g) * 3).to(dtype)
    w = (1 + 0.2 * torch.randn(C, device=dev, generator=g)).to(dtype)
    b = (0.3 * torch.randn(C, device=dev, generator=g)).to(dtype)
    y = ops.layer_norm(x, w, b, 1e-5)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w.float(), b.float(), 1e-5)
    assert y.dtype == dtype and y.shape == x.shape
    assert float((y.float() - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    lib = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
    assert float((y.float() - lib.float()).abs().max()) <= tol * max(1.0, float(ref.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_phase_shuffle_bias_nhwc_bit_exact(dtype):
    """rd_phase_shuffle_bias_nhwc against its index formula (include/rd_b200.h) evaluated by torch: a pure gather plus one
    fp32 add and rounding -> identical bits; with and without bias, odd sizes, one-pixel input."""
    from rollingdepth_b200 import ops
    from test_models_cpu import phase_shuffle_reference
    dev = _dev()
    g = torch.Generator().manual_seed(9)
    for (N, C, H, W) in [(2, 64, 5, 7), (1, 8, 1, 1), (3, 320, 13, 6), (1, 128, 54, 96)]:
        t = torch.randn(N, 4 * C, H + 1, W + 1, generator=g).to(dtype).to(dev).contiguous(memory_format=torch.channels_last)
        b = torch.randn(C, generator=g).to(dtype).to(dev)
        for bias in (b, None):
            y = ops.phase_shuffle_bias_nhwc(t, bias, H, W)
            assert y.shape == (N, C, 2 * H, 2 * W) and y.is_contiguous(memory_format=torch.channels_last)
            assert torch.equal(y, phase_shuffle_reference(t, bias, H, W))
    with pytest.raises(Exception):
        ops.phase_shuffle_bias_nhwc(t, None, H + 1, W)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.float16, 3e-3), (torch.bfloat16, 2.5e-2)])
def test_upsample2d_subpixel_matches_upsample_then_conv(dtype, tol):
    """Upsample2D in sub-pixel form (2x2 phase convolution + rd_phase_shuffle_bias_nhwc) against the reference's order of
    operations (upsampling.py Upsample2D.forward: F.interpolate nearest, then conv) evaluated by torch in fp32 on the same
    weights: relative to the output range, within the rounding of the dtype; explicit non-2x sizes keep the nearest + conv path."""
    from rollingdepth_b200.sd2.layers import Upsample2D
    dev = _dev()
    torch.manual_seed(5)
    torch.backends.cudnn.allow_tf32 = False          # torch allows TF32 convolutions by default: the fp32 case means fp32
    for (N, C, H, W) in [(2, 64, 9, 12), (1, 320, 14, 24), (3, 128, 27, 48)]:
        m = Upsample2D(C).to(dev, dtype).to(memory_format=torch.channels_last).eval()
        x = torch.randn(N, C, H, W, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            ref = torch.nn.functional.conv2d(torch.nn.functional.interpolate(x.float(), scale_factor=2.0, mode="nearest"),
                                             m.conv.weight.float(), m.conv.bias.float(), 1, 1)
            m.fused = True
            m.conv.fused = True
            got = m(x)
            assert got.shape == ref.shape and got.dtype == dtype
            scale = float(ref.abs().max())
            assert float((got.float() - ref).abs().max()) <= tol * scale
            m.subpixel = False
            old = m(x)
            assert float((old.float() - ref).abs().max()) <= tol * scale
            odd = m(x, output_size=(2 * H - 1, 2 * W - 1))          # UNet levels of odd size pass an explicit size
            m.subpixel = True
            assert torch.equal(odd, m(x, output_size=(2 * H - 1, 2 * W - 1)))
            # weights changed in place -> the cached phase kernels are rebuilt
            m.conv.weight.mul_(2.0)
            got2 = m(x)
            ref2 = torch.nn.functional.conv2d(torch.nn.functional.interpolate(x.float(), scale_factor=2.0, mode="nearest"),
                                              m.conv.weight.float(), m.conv.bias.float(), 1, 1)
            assert float((got2.float() - ref2).abs().max()) <= tol * float(ref2.abs().max())
