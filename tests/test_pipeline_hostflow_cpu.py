"""The product pipeline's HOST control flow on CPU against the REFERENCE pipeline's own outputs.

`RollingDepthPipeline.init_snippet_infer` (dilation groups, flattened snippet list, balanced batches of several snippets per
UNet call, decode straight into the gather buffer), `refine` (annealed windows, batched, `_average_windows`), `encode_rgb` /
`decode_depth` and `B200XFrameAttnProcessor` are host code around four calls into librd_b200.so.  Here those calls
(`ops.xattn_fwd`, `ops.ddim_step`, `ops.ddim_add_noise`; the aligner as a whole) are replaced by the oracle's CPU arithmetic --
test-only: the product has no CPU path and `forward()` itself refuses a CPU device -- and the stages are chained exactly as
`forward` chains them.  The results are compared with tests/golden/models_pipeline.npz, i.e. with what the reference's
`forward` (rollingdepth/rollingdepth_pipeline.py:194-353, one snippet per UNet call) produced for the same frames, weights,
text embedding and noise generator.  The kernels themselves are pinned by the -m gpu tier.
"""
import os

import numpy as np
import pytest
import torch

import model_utils as mu
from oracle import coalign as oc
from oracle import xframe_attention as oxa


@pytest.fixture()
def library_calls_stubbed_by_the_oracle(monkeypatch):
    from rollingdepth_b200 import _lib as L
    from rollingdepth_b200 import ops
    calls = dict(xattn=[], ddim_step=0, add_noise=0)

    def xattn(q, k, v, scale=None, out=None, flags=0):
        calls["xattn"].append(tuple(q.shape))
        o = oxa.sdpa_ref(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float(), scale=scale)
        return o.transpose(1, 2).to(q.dtype)

    def ddim_step(model_output, sample, sa, sb, sap, sbp, prediction_type, clip=0.0, out=None):
        # scheduling_ddim.py:405-452 (eta = 0) with the host-side scalars, v-prediction only (RollingDepth's scheduler)
        assert prediction_type == "v_prediction" and clip == 0.0
        sa, sb, sap, sbp = [torch.tensor(s, dtype=torch.float32) for s in (sa, sb, sap, sbp)]
        x0 = sa * sample - sb * model_output
        eps = sa * model_output + sb * sample
        y = sap * x0 + sbp * eps
        calls["ddim_step"] += 1
        if out is None:
            return y
        out.copy_(y)
        return out

    def add_noise(original, noise, sa, sb):
        calls["add_noise"] += 1
        return torch.tensor(sa, dtype=original.dtype) * original + torch.tensor(sb, dtype=original.dtype) * noise

    monkeypatch.setattr(ops, "xattn_fwd", xattn)
    monkeypatch.setattr(ops, "ddim_step", ddim_step)
    monkeypatch.setattr(ops, "ddim_add_noise", add_noise)
    monkeypatch.setattr(L, "require_cuda", lambda *a, **k: None)
    return calls


def _absrel(a, b):
    d0, d1 = (a.float() + 1) / 2, (b.float() + 1) / 2
    return float(((d0 - d1).abs() / d1.clamp_min(1e-3)).mean())


@pytest.mark.parametrize("tag,dil,refine", [("norefine", [1, 3], 0), ("refine", [1, 2, 3], 2)])
@pytest.mark.parametrize("snippet_batch,max_vae_bs", [(1, 4), (4, 4), (16, 36)])
def test_host_control_flow_reproduces_the_reference_run(golden_dir, library_calls_stubbed_by_the_oracle, tag, dil, refine,
                                                        snippet_batch, max_vae_bs):
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    from rollingdepth_b200.sd2.unet import UNet2DConditionModel
    from rollingdepth_b200.sd2.vae import AutoencoderKL
    calls = library_calls_stubbed_by_the_oracle
    g = np.load(os.path.join(golden_dir, "models_pipeline.npz"))
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=snippet_batch)     # installs the processor
    pipe.empty_text_embed = torch.from_numpy(g["pipe_text"])
    frames = torch.from_numpy(g["pipe_frames"])
    N = frames.shape[1]
    with torch.no_grad():
        # ---- the stages of forward(), in forward()'s order (rollingdepth_pipeline.py:244-345)
        dil = [pipe.cap_max_dilation(seq_len=N, snippet_len=3, dilation=d, verbose=False) for d in dil]
        start = pipe.cap_max_dilation(seq_len=N, snippet_len=3, dilation=3, verbose=False)
        rgb_latent = pipe.encode_rgb(frames, max_batch_size=max_vae_bs)
        _, _, _, h, w = rgb_latent.shape
        noise = torch.randn((1, 4, h, w), generator=torch.Generator().manual_seed(0))
        init_noise = noise[:, None].expand(1, N, 4, h, w)
        snips = pipe.init_snippet_infer(rgb_latent=rgb_latent, init_noise=init_noise, dilations=dil, cap_dilation=True,
                                        snippet_lengths=[3] * len(dil), init_infer_steps=[1] * len(dil), strides=[1] * len(dil),
                                        max_vae_bs=max_vae_bs, unload_snippet=False, verbose=False)
        assert len(snips) == len(dil)
        for i, s in enumerate(snips):
            ref = torch.from_numpy(g[f"pipe_{tag}_snip{i}"])
            assert s.shape == ref.shape and float((s - ref).abs().max()) < 2e-5, (i, float((s - ref).abs().max()))
        # cross-frame attention really ran over (snippets per call) x (3 frames x tokens): the batched call is the product's
        n_snip = sum(N - 2 * d for d in dil)
        n_calls = -(-n_snip // snippet_batch)                       # all dilations share (length, steps): ONE flattened list
        assert calls["ddim_step"] == n_calls                        # one UNet call + one DDIM step per balanced batch
        assert max(shape[0] for shape in calls["xattn"]) == -(-n_snip // n_calls)       # e.g. 20 snippets, batch 16 -> 10 + 10
        coal, _, _, _ = oc.DepthAlignerRef(num_iterations=30).run(snips, dil)
        coal = oc.renorm(coal)
        # the aligner (closed-form oracle, 30 sign-gradient Adam steps) amplifies 1e-6 snippet differences: same tolerance as the
        # oracle-pipeline test and the north star (1e-3); measured 1.3e-4
        assert _absrel(coal, torch.from_numpy(g[f"pipe_{tag}_coaligned"])) <= 1e-3
        if refine:
            dl = pipe.encode_rgb(coal[None].expand(1, N, 3, *coal.shape[-2:]), max_batch_size=max_vae_bs)
            new = pipe.refine(rgb_latent=rgb_latent, depth_latents=dl, init_noise=init_noise, refine_step=refine, snippet_len=3,
                              start_dilation=start, verbose=False)
            assert calls["add_noise"] == 1
            depth_pred = pipe.decode_depth(new, max_batch_size=max_vae_bs).squeeze(0)
        else:
            depth_pred = coal
    ref_pred = torch.from_numpy(g[f"pipe_{tag}_depth_pred"])
    assert depth_pred.shape == ref_pred.shape
    assert _absrel(depth_pred, ref_pred) <= 1e-3          # north-star tolerance


_SHARDED_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["RD_ROOT"]); sys.path.insert(0, os.path.join(os.environ["RD_ROOT"], "tests"))
import model_utils as mu
from oracle import xframe_attention as oxa
from rollingdepth_b200 import _lib as L, ops
from rollingdepth_b200.pipeline import RollingDepthPipeline
from rollingdepth_b200.scheduler import DDIMSchedule
from rollingdepth_b200.sd2.unet import UNet2DConditionModel
from rollingdepth_b200.sd2.vae import AutoencoderKL

# the four library calls, stubbed by the oracle's CPU arithmetic (test-only; see the module docstring)
def xattn(q, k, v, scale=None, out=None, flags=0):
    o = oxa.sdpa_ref(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float(), scale=scale)
    return o.transpose(1, 2).to(q.dtype)
def ddim_step(mo, x, sa, sb, sap, sbp, ptype, clip=0.0, out=None):
    sa, sb, sap, sbp = [torch.tensor(s, dtype=torch.float32) for s in (sa, sb, sap, sbp)]
    y = sap * (sa * x - sb * mo) + sbp * (sa * mo + sb * x)
    if out is None: return y
    out.copy_(y); return out
def add_noise(x, n, sa, sb):
    return torch.tensor(sa, dtype=x.dtype) * x + torch.tensor(sb, dtype=x.dtype) * n
ops.xattn_fwd, ops.ddim_step, ops.ddim_add_noise = xattn, ddim_step, add_noise
L.require_cuda = lambda *a, **k: None

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = np.load(os.path.join(os.environ["RD_ROOT"], "tests", "golden", "models_pipeline.npz"))
unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
frames = torch.from_numpy(g["pipe_frames"]); N = frames.shape[1]
text = torch.from_numpy(g["pipe_text"])

def stages(pipe, dil, refine):
    # snippet_batch = max_vae_bs = 1 on both sides: oneDNN's bits depend on the batch size, the partition must not
    with torch.no_grad():
        dil = [pipe.cap_max_dilation(seq_len=N, snippet_len=3, dilation=d, verbose=False) for d in dil]
        start = pipe.cap_max_dilation(seq_len=N, snippet_len=3, dilation=3, verbose=False)
        rgb_latent = pipe.encode_rgb(frames, max_batch_size=1, sharded=True)
        h, w = rgb_latent.shape[-2:]
        noise = torch.randn((1, 4, h, w), generator=torch.Generator().manual_seed(0))
        pipe.shard.broadcast_(noise)
        init_noise = noise[:, None].expand(1, N, 4, h, w)
        snips = pipe.init_snippet_infer(rgb_latent=rgb_latent, init_noise=init_noise, dilations=dil, cap_dilation=True,
                                        snippet_lengths=[3] * len(dil), init_infer_steps=[1] * len(dil), strides=[1] * len(dil),
                                        max_vae_bs=1, unload_snippet=False, verbose=False)
        new = None
        if refine:
            depth = torch.stack([s[:, 0] for s in snips[:1]])[0][:N]          # any [n, 1, H, W] stand-in for the aligned depth
            depth = torch.cat([depth, depth[-1:].expand(N - depth.shape[0], -1, -1, -1)])
            dl = pipe.encode_rgb(depth[None].expand(1, N, 3, *depth.shape[-2:]), max_batch_size=1, sharded=True)
            new = pipe.refine(rgb_latent=rgb_latent, depth_latents=dl, init_noise=init_noise, refine_step=refine, snippet_len=3,
                              start_dilation=start, verbose=False)
            new = pipe.decode_depth(new, max_batch_size=1, sharded=True)
    return snips, new

for tag, dil, refine in (("norefine", [1, 3], 0), ("refine", [1, 2, 3], 2)):
    single = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=1)
    single.empty_text_embed = text
    s1, r1 = stages(single, dil, refine)
    sharded = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=1)
    sharded.empty_text_embed = text
    sharded.enable_sharding(output_ranks=None)
    assert sharded.shard.world == world and sharded.shard.rank == rank
    s2, r2 = stages(sharded, dil, refine)
    assert len(s1) == len(s2)
    for i, (a, b) in enumerate(zip(s1, s2)):          # every rank holds the COMPLETE snippet list, identical to one rank's
        assert a.shape == b.shape and torch.equal(a, b), (tag, i, float((a - b).abs().max()))
        ref = torch.from_numpy(g[f"pipe_{tag}_snip{i}"])
        assert float((b - ref).abs().max()) < 2e-5     # ... and equal to the reference pipeline's snippets
    if refine:
        assert torch.equal(r1, r2), float((r1 - r2).abs().max())
dist.destroy_process_group()
print("rank ok")
"""


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_host_control_flow_gloo(tmp_path, world):
    """The snippet-sharded path on CPU (gloo, `world` processes): frame-sharded VAE encode, the flattened (dilation, snippet)
    list block-partitioned over the ranks and exchanged with one in-place all-gather, refinement windows sharded per step,
    frame-sharded final decode -- with the library calls stubbed by the oracle.  Every rank ends with the complete snippet
    list, bit-identical to the single-rank run of the same code and equal to the REFERENCE pipeline's snippets (fixture);
    the refined and decoded result is bit-identical to the single-rank one.  10 frames over 3 ranks exercises short and
    empty trailing blocks (8 + 6 snippets -> 5 + 5 + 4; refinement windows 6 -> 2 + 2 + 2 / 8 -> 3 + 3 + 2)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_SHARDED_WORKER)
    env = dict(os.environ, RD_ROOT=root, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), str(script)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("rank ok") == world
