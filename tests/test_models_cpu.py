"""CPU checks of the host-side model code and of the oracle pipeline against reference fixtures
(tests/golden/models_pipeline.npz, sd2_state_keys.json -- produced by oracle/gen_golden_pipeline.py from the
reference's vendored diffusers + rollingdepth_pipeline.py)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import model_utils as mu
from oracle import pipeline as opipe
from rollingdepth_b200.sd2.unet import UNet2DConditionModel
from rollingdepth_b200.sd2.vae import AutoencoderKL


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "models_pipeline.npz"))


def _strip(cfg):
    return {k: v for k, v in cfg.items() if k not in ("down_block_types", "up_block_types") or "Encoder" not in str(v)}


def test_state_dict_keys_match_diffusers(golden_dir):
    """A diffusers SD2 checkpoint (e.g. prs-eth/rollingdepth-v1-0) must load with strict=True."""
    ref = json.load(open(os.path.join(golden_dir, "sd2_state_keys.json")))
    with torch.device("meta"):
        unet = UNet2DConditionModel(**mu.SD2_UNET)
        vae = AutoencoderKL(**{k: v for k, v in mu.SD_VAE.items() if not k.endswith("block_types")})
    ours_u = {k: list(v.shape) for k, v in unet.state_dict().items()}
    ours_v = {k: list(v.shape) for k, v in vae.state_dict().items()}
    assert ours_u == ref["unet"]
    assert ours_v == ref["vae"]
    assert 860e6 < sum(int(np.prod(s)) for s in ours_u.values()) < 870e6   # 866 M parameters


def _tiny():
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
    return unet, vae


@pytest.mark.parametrize("tag", ["even", "odd", "n4"])
def test_unet_forward_matches_reference(gold, tag):
    unet, _ = _tiny()
    x = torch.from_numpy(gold[f"unet_{tag}_x"])
    ctx = torch.from_numpy(gold[f"unet_{tag}_ctx"])
    n = x.shape[0]
    with torch.no_grad():
        y = unet(x, torch.full((n,), 999), encoder_hidden_states=ctx, num_view=n).sample
        y_pf = unet(x, torch.full((n,), 999), encoder_hidden_states=ctx, num_view=None).sample
    assert float((y - torch.from_numpy(gold[f"unet_{tag}_y"])).abs().max()) < 2e-5
    assert float((y_pf - torch.from_numpy(gold[f"unet_{tag}_y_perframe"])).abs().max()) < 2e-5
    assert float((y - y_pf).abs().max()) > 1e-4   # cross-frame attention must matter


def test_vae_matches_reference(gold):
    _, vae = _tiny()
    img = torch.from_numpy(gold["vae_img"])
    with torch.no_grad():
        mom = vae.quant_conv(vae.encoder(img))
        dec = vae.decoder(vae.post_quant_conv(mom[:, :4]))
    assert float((mom - torch.from_numpy(gold["vae_moments"])).abs().max()) < 2e-5
    assert float((dec - torch.from_numpy(gold["vae_dec"])).abs().max()) < 2e-5


def test_attention_refuses_processor_without_num_view():
    from rollingdepth_b200.sd2.layers import Attention

    class NoView:
        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
            return hidden_states

    a = Attention(64, None, 1, 64, processor=NoView())
    with pytest.raises(TypeError):
        a(torch.zeros(3, 4, 64), num_view=3)


@pytest.mark.parametrize("tag,kw", [("norefine", dict(dilations=[1, 3], refine_step=0)),
                                    ("refine", dict(dilations=[1, 2, 3], refine_step=2))])
def test_oracle_pipeline_matches_reference(gold, tag, kw):
    """The oracle's restatement of RollingDepthPipeline.forward against the reference's own output
    (CPU fp32, tiny SD2-architecture models, 10 frames)."""
    unet, vae = _tiny()
    frames = torch.from_numpy(gold["pipe_frames"])
    res = opipe.forward(unet, vae, mu.DDIM_CFG, torch.from_numpy(gold["pipe_text"]), frames, cap_dilation=True,
                        coalign_kwargs=dict(num_iterations=30), refine_start_dilation=3,
                        generator=torch.Generator().manual_seed(0), **kw)
    for i, s in enumerate(res.snippet_ls):
        assert float((s - torch.from_numpy(gold[f"pipe_{tag}_snip{i}"])).abs().max()) < 5e-5
    co = torch.from_numpy(gold[f"pipe_{tag}_coaligned"])
    dp = torch.from_numpy(gold[f"pipe_{tag}_depth_pred"])
    assert float((res.depth_coaligned.squeeze(0) - co).abs().max()) < 2e-3
    assert float((res.depth_pred - dp).abs().max()) < 5e-3
    d0, d1 = (res.depth_pred + 1) / 2, (dp + 1) / 2
    absrel = float(((d0 - d1).abs() / d1.clamp_min(1e-3)).mean())
    assert absrel <= 1e-3, absrel


def test_unsupported_config_values_are_refused():
    """A diffusers config.json key these modules do not implement must raise instead of loading with strict=True and
    silently computing something else (only the values the SD2 / RollingDepth checkpoints use are accepted)."""
    with torch.device("meta"):
        UNet2DConditionModel(**mu.SD2_UNET, upcast_attention=True, sample_size=96, act_fn="silu", _class_name="UNet2DConditionModel")
        AutoencoderKL(**mu.SD_VAE, sample_size=768, force_upcast=True)
        for bad in (dict(only_cross_attention=True), dict(resnet_time_scale_shift="scale_shift"), dict(norm_eps=1e-6),
                    dict(transformer_layers_per_block=2), dict(class_embed_type="timestep"), dict(not_a_key=1)):
            with pytest.raises(NotImplementedError):
                UNet2DConditionModel(**mu.SD2_UNET, **bad)
        with pytest.raises(NotImplementedError):
            AutoencoderKL(**{k: v for k, v in mu.SD_VAE.items() if not k.endswith("block_types")}, mid_block_add_attention=False)


@pytest.mark.parametrize("H,W", [(40, 48), (50, 70), (37, 91), (432, 768)])
def test_latent_size_formula_matches_the_encoder(H, W):
    """`RollingDepthPipeline._latent_hw` (used to size the all-gather buffers, also on ranks that own no frame) against
    the shape the VAE encoder really produces, including sizes that are not multiples of the total stride: every level
    pads (0, 1, 0, 1) and applies a 3x3 stride-2 convolution without padding = floor(size / 2)."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    unet, vae = _tiny()
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), install_processor=False)
    if H * W < 10000:
        with torch.no_grad():
            lat = vae.quant_conv(vae.encoder(torch.zeros(1, 3, H, W)))
        assert tuple(lat.shape[-2:]) == pipe._latent_hw(H, W)
        dec_hw = pipe._decoded_hw(*lat.shape[-2:])
        with torch.no_grad():
            dec = vae.decoder(vae.post_quant_conv(lat[:, :4]))
        assert tuple(dec.shape[-2:]) == dec_hw
    else:
        assert pipe._latent_hw(H, W) == (H // 2, W // 2)        # the tiny VAE has one stride-2 level


def phase_shuffle_reference(t, bias, H, W):
    """torch restatement of rd_phase_shuffle_bias_nhwc (include/rd_b200.h)."""
    C = t.shape[1] // 4
    y = t.new_empty(t.shape[0], C, 2 * H, 2 * W)
    for a in range(2):
        for b in range(2):
            y[:, :, a::2, b::2] = t[:, (2 * a + b) * C:(2 * a + b + 1) * C, a:a + H, b:b + W]
    if bias is not None:
        y = (y.float() + bias.float().view(1, -1, 1, 1)).to(t.dtype) if t.dtype != torch.float64 else y + bias.view(1, -1, 1, 1)
    return y


@pytest.mark.parametrize("H,W", [(1, 1), (6, 9), (7, 4)])
def test_subpixel_upsample_convolution_is_the_same_linear_map(H, W):
    """Upsample2D in sub-pixel form (one 2x2 convolution with the four phase kernels + phase interleave) against
    F.interpolate(nearest, 2x) + the 3x3 convolution (upsampling.py Upsample2D.forward) in float64: identical up to
    summation order, borders included."""
    from rollingdepth_b200.sd2.layers import subpixel_phase_weights
    g = torch.Generator().manual_seed(H * 31 + W)
    x = torch.randn(2, 5, H, W, dtype=torch.float64, generator=g)
    w = torch.randn(7, 5, 3, 3, dtype=torch.float64, generator=g)
    bias = torch.randn(7, dtype=torch.float64, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, bias, 1, 1)
    pw = subpixel_phase_weights(w)
    assert pw.shape == (28, 5, 2, 2)
    got = phase_shuffle_reference(F.conv2d(x, pw, None, 1, 1), bias, H, W)
    assert float((got - ref).abs().max()) < 1e-12


_ENCODE_WORKER = r"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["RD_ROOT"]); sys.path.insert(0, os.path.join(os.environ["RD_ROOT"], "tests"))
import model_utils as mu
from rollingdepth_b200.pipeline import RollingDepthPipeline
from rollingdepth_b200.scheduler import DDIMSchedule
from rollingdepth_b200.sd2.unet import UNet2DConditionModel
from rollingdepth_b200.sd2.vae import AutoencoderKL
dist.init_process_group("gloo")
unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), install_processor=False)
with torch.no_grad():
    for n_frames, H, W in ((1, 37, 51), (2, 40, 45), (5, 33, 32)):     # fewer frames than ranks; sizes that are not multiples of the stride
        rgb = torch.rand(1, n_frames, 3, H, W, generator=torch.Generator().manual_seed(n_frames)) * 2 - 1
        pipe.shard = type(pipe.shard)()                                 # one rank
        ref = pipe.encode_rgb(rgb, 1)      # one frame per VAE call on every path: oneDNN's bits depend on the batch size
        pipe.enable_sharding(output_ranks=None)
        got = pipe.encode_rgb(rgb, 1, sharded=True)
        assert got.shape == ref.shape and torch.equal(got, ref), (n_frames, H, W, got.shape, ref.shape)
        lo, hi = pipe.shard.block(n_frames)                             # this rank uploads only its own block of frames
        got2 = pipe.encode_rgb(rgb[:, lo:hi], 1, sharded=True, frame_block=(lo, hi), total_frames=n_frames)
        assert torch.equal(got2, ref)
        lat = ref[:, :, :4]
        ref_d = pipe.decode_depth(lat, 1)
        got_d = pipe.decode_depth(lat, 1, sharded=True)
        assert torch.equal(got_d, ref_d)
dist.destroy_process_group()
print("rank ok")
"""


def test_sharded_vae_stages_world3_gloo_odd_sizes(tmp_path):
    """frame-sharded VAE encode / decode with more ranks than frames and frame sizes that are not multiples of
    the VAE stride -- every rank must size its gather slots identically (`_latent_hw`, `_decoded_hw`) and the gathered
    result must equal the single-rank one bit for bit (gloo, CPU modules)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_ENCODE_WORKER)
    env = dict(os.environ, RD_ROOT=root, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("rank ok") == 3


def test_oracle_chain_at_full_sd2_architecture_matches_reference_cfg1(golden_dir):
    """BASELINE.json configs[0] on CPU fp32 with the FULL SD2 architecture (866 M UNet + 84 M VAE, name-keyed weights):
    the oracle's encode -> UNet step (cross-frame attention, num_view = 3) -> DDIM -> decode chain, built from this repo's
    `sd2` modules, against the decoded snippets of the REFERENCE pipeline's own run (tests/golden/cfg1_reference.npz) for the
    five snippets the fixture stores (both dilations).  This closes the two-hop chain at full size: reference == oracle here,
    oracle ~ product on the B200 (`test_real_size_fp16_matches_oracle_on_device`); the aligner / merge / renorm half of the
    oracle is pinned bit for bit by tests/test_oracle_vs_golden.py."""
    from oracle import schedule as osch
    g = np.load(os.path.join(golden_dir, "cfg1_reference.npz"))
    torch.manual_seed(0)
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.SD2_UNET).eval(), seed=11, gain=0.7)
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.SD_VAE.items() if not k.endswith("block_types")}).eval(),
                                 seed=12, gain=0.7)
    frames = mu.synth_video(16, 256, 256, seed=0)
    text = torch.from_numpy(g["text"])
    sch = osch.DDIMRef(**mu.DDIM_CFG)
    ts = sch.set_timesteps(1)
    worst = 0.0
    with torch.no_grad():
        rgb_latent = opipe.encode_rgb(vae, frames, 4)
        noise = torch.randn((1, 4, 32, 32), generator=torch.Generator().manual_seed(0))[:, None].expand(1, 16, 4, 32, 32)
        for gi, d in enumerate([1, 4]):
            idx_ls = osch.get_snippet_indice(0, 1, 16, 3, d, d, 1)
            assert len(idx_ls) == int(g[f"snip{gi}_shape"][0])
            rows = sorted(int(k.split("_")[1]) for k in g.files if k.startswith(f"snip{gi}_") and not k.endswith("shape"))
            for r in rows:
                idx = idx_ls[r]
                dep = noise[:, idx].clone()
                pred = opipe.single_step(unet, rgb_latent[:, idx], dep, int(ts[0]), text)
                dec = opipe.decode_depth(vae, sch.step(pred, int(ts[0]), dep), 4)[0]          # [3, 1, 256, 256]
                ref = torch.from_numpy(g[f"snip{gi}_{r}"])
                worst = max(worst, float((dec[:, :, ::2, ::2] - ref).abs().max()) / max(1.0, float(ref.abs().max())))
    assert worst < 5e-5, worst      # fp32 on both sides; measured 9e-6 of the range (max-abs 1.1e-5: conv / GEMM blocking only)


def test_forward_argument_handling_matches_reference(golden_dir):
    """`RollingDepthPipeline.forward` argument checks and in-place dilation capping (rollingdepth_pipeline.py:213-258)
    against what the REFERENCE's forward does with the same calls (tests/golden/forward_args.json, recorded by
    oracle/gen_golden.py::gen_forward_args): rejected calls raise the same exception class with the same message; accepted
    calls leave the caller's `dilations` list capped in place exactly like the reference (Q6) -- and then, on this CPU-resident
    pipeline, stop with RdError because there is no CPU fallback."""
    import copy
    from oracle.gen_golden import FORWARD_ARG_DEFAULTS
    from rollingdepth_b200._lib import RdError
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    unet, vae = _tiny()
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), install_processor=False)
    cases = json.load(open(os.path.join(golden_dir, "forward_args.json")))
    assert len(cases) >= 15 and sum(not c["accepted"] for c in cases) >= 9
    for c in cases:
        args = copy.deepcopy(dict(FORWARD_ARG_DEFAULTS, **c["kwargs"]))
        frames = torch.zeros(1, c["seq_len"], 3, 16, 16)
        if c["accepted"]:
            with pytest.raises(RdError, match="no CPU fallback"):
                pipe.forward(input_frames=frames, **args)
            assert list(args["dilations"]) == c["dilations_after"], c
        else:
            exc = {"AssertionError": AssertionError, "NotImplementedError": NotImplementedError}[c["error"]]
            with pytest.raises(exc) as ei:
                pipe.forward(input_frames=frames, **args)
            assert str(ei.value) == c["message"], c


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_refine_window_average_is_bit_identical_to_the_reference_loop(dtype):
    """`RollingDepthPipeline._average_windows` (vectorised over the windows of one refinement step) against the reference's
    formulation (rollingdepth_pipeline.py:585-629: `new *= 0`, then for every window in ascending order `new[:, idx] += pred`
    and `count[:, idx] += 1`, then `new / count` with an int32 count): identical bits in fp16 / bf16 / fp32, because the
    per-frame order of the in-dtype additions is the same (ascending window start = descending position in the window)."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline as P
    g0 = torch.Generator().manual_seed(7)
    for seq_len, n, start, total in ((16, 3, 5, 10), (40, 3, 6, 5), (30, 4, 4, 7), (12, 2, 6, 4), (9, 1, 1, 2)):
        for i_step in range(total):
            idx_ls = P.get_snippet_indice(i_step, [0] * total, seq_len, n, start if n > 1 else 1, 1, 1)
            n_win = len(idx_ls)
            gap = idx_ls[0][1] - idx_ls[0][0] if n > 1 else 1
            preds = (torch.randn(n_win, n, 4, 3, 5, generator=g0) * 3).to(dtype)
            old = torch.randn(1, seq_len, 4, 3, 5, generator=g0).to(dtype)
            new = old.clone()
            new *= 0.0
            count = torch.zeros((1, seq_len, 1, 1, 1)).int()
            for k, idx in enumerate(idx_ls):
                new[:, idx, :, :, :] += preds[k][None].clone()
                count[:, idx, 0, 0, 0] += 1
            if not bool(torch.all(count > 0)):
                with pytest.raises(AssertionError):
                    P._average_windows(preds, gap, seq_len)
                continue
            ref = new / count
            got = P._average_windows(preds, gap, seq_len)[None]
            assert got.dtype == ref.dtype == dtype and torch.equal(got, ref), (seq_len, n, start, i_step)


def test_product_vae_wrappers_and_single_step_match_the_oracle_on_cpu():
    """`RollingDepthPipeline.encode_rgb / decode_depth / single_step` are plain host code around the modules (no kernel of the
    library): on CPU they must agree with the oracle's restatement of rollingdepth_pipeline.py:635-740 -- posterior mean (not a
    sample), the 0.18215 factors, the mean over the decoder's three channels, frame order across chunks -- for chunk sizes that
    divide the frame count and for ones that do not (the product balances its chunks, the reference takes fixed ones: same
    frames, other batch compositions, so equality is to fp32 conv round-off, not bit for bit)."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    unet, vae = _tiny()
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), install_processor=False)
    g = torch.Generator().manual_seed(3)
    frames = torch.rand(1, 7, 3, 24, 40, generator=g) * 2 - 1
    with torch.no_grad():
        ref_lat = opipe.encode_rgb(vae, frames, 4)
        for bs in (1, 3, 4, 7, 36):
            lat = pipe.encode_rgb(frames, max_batch_size=bs)
            assert lat.shape == ref_lat.shape == (1, 7, 4, 12, 20)
            assert float((lat - ref_lat).abs().max()) < 1e-5, bs
        snip = ref_lat[:, :6].reshape(2, 3, 4, 12, 20)                        # two "snippets" of three frames
        ref_dec = opipe.decode_depth(vae, snip, 4)
        for bs in (1, 4, 5, 36):
            dec = pipe.decode_depth(snip, max_batch_size=bs)
            assert dec.shape == ref_dec.shape == (2, 3, 1, 24, 40)
            assert float((dec - ref_dec).abs().max()) < 1e-5, bs
        out = torch.empty(2, 3, 1, 24, 40)
        assert pipe.decode_depth(snip, max_batch_size=4, out=out) is out and float((out - ref_dec).abs().max()) < 1e-5
        # one UNet call over b snippets at once == the reference's one call per snippet (num_view rearrange is per batch element)
        text = torch.randn(1, 2, 32, generator=g)
        rgb, dep = snip, torch.randn(2, 3, 4, 12, 20, generator=g)
        both = pipe.single_step(rgb, dep, 999, text)
        for b in range(2):
            one = opipe.single_step(unet, rgb[b:b + 1], dep[b:b + 1], 999, text)
            assert float((both[b:b + 1] - one).abs().max()) < 2e-5
