"""Generate tests/golden/* by running the REFERENCE's own code (build container only).

    python oracle/gen_golden.py [--only snippets,snippets_wide,forward_args,aligner_geometry,coalign,xattn,ddim,ddim_wide,pipeline,cfg1,cfg1_loss_floor]

/root/reference is imported read-only with the recipe of SURVEY.md App. C (the
`rollingdepth` package __init__ is bypassed, PyAV is stubbed, one transformers
constant is shimmed).  The fixtures are small and are committed together with
this script; nothing at test / bench time reads /root/reference.
"""
from __future__ import annotations

import argparse
import contextlib
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def import_reference():
    import transformers.utils as tu
    tu.FLAX_WEIGHTS_NAME = "flax_model.msgpack"                       # shim 1
    sys.modules.setdefault("av", types.ModuleType("av"))              # shim 2
    sys.path.insert(0, f"{REF}/diffusers/src")
    sys.dont_write_bytecode = True
    import diffusers  # noqa: F401
    pkg = types.ModuleType("rollingdepth")
    pkg.__path__ = [f"{REF}/rollingdepth"]
    sys.modules["rollingdepth"] = pkg                                 # detour around __init__

    def load(n):
        s = importlib.util.spec_from_file_location(f"rollingdepth.{n}", f"{REF}/rollingdepth/{n}.py")
        m = importlib.util.module_from_spec(s)
        sys.modules[f"rollingdepth.{n}"] = m
        s.loader.exec_module(m)
        return m

    load("video_io")
    da = load("depth_aligner")
    rp = load("rollingdepth_pipeline")
    return da, rp


# --------------------------------------------------------------------------- snippets
def gen_snippets(da, rp):
    P = rp.RollingDepthPipeline
    cases = []
    for seq_len in (7, 16, 30, 150):
        for n in (2, 3, 4):
            for d in (1, 4, 10, 25):
                dc = P.cap_max_dilation(seq_len=seq_len, snippet_len=n, dilation=d, verbose=False)
                if dc < 1:
                    continue
                idx = P.get_snippet_indice(0, torch.zeros(1), seq_len, n, dc, dc, 1)
                cases.append(dict(kind="init", seq_len=seq_len, snippet_len=n, dilation=d, capped=dc, idx=idx))
    for seq_len, start, total in ((16, 5, 10), (150, 6, 10), (40, 6, 5), (150, 6, 20), (30, 4, 7)):
        for i_step in range(total):
            idx = P.get_snippet_indice(i_step, torch.zeros(total), seq_len, 3, start, 1, 1)
            cases.append(dict(kind="refine", seq_len=seq_len, snippet_len=3, start=start, total=total,
                              i_step=i_step, idx=idx))
    with open(os.path.join(GOLD, "snippets.json"), "w") as f:
        json.dump(cases, f)
    print("snippets:", len(cases), "cases")


def snippet_digest(idx):
    """Compact, exact fingerprint of a window index list (the wide fixture stores these instead of the lists)."""
    import zlib
    flat = [x for win in idx for x in win]
    return dict(n=len(idx), first=idx[0], last=idx[-1], index_sum=sum(flat),
                crc=zlib.crc32(",".join(map(str, flat)).encode()) & 0xFFFFFFFF)


def gen_snippets_wide(da, rp):
    """A wider sweep of the reference's scheduler than snippets.json (strides 1-4, sequences up to 1000 frames, snippet lengths
    2-8, annealed dilations with several start / end pairs, every step) stored as digests, plus the reference's behaviour on
    invalid requests (exception class).  tests/golden/snippets_wide.json."""
    import logging
    logging.disable(logging.WARNING)            # "Not every frame is covered" for strides > 1: expected
    P = rp.RollingDepthPipeline
    cases = []

    def one(**a):
        try:
            idx = P.get_snippet_indice(a["i_step"], torch.zeros(a["total"]), a["seq_len"], a["snippet_len"], a["dilation_start"],
                                       a["dilation_end"], a["stride"])
            cases.append(dict(a, **snippet_digest(idx)))
        except Exception as e:  # noqa: BLE001
            cases.append(dict(a, error=type(e).__name__))

    for seq_len in (3, 5, 9, 16, 31, 150, 1000):
        for n in (2, 3, 5, 8):
            for d in (1, 2, 7, 25, 60):
                for stride in (1, 2, 3, 4):
                    one(i_step=0, total=1, seq_len=seq_len, snippet_len=n, dilation_start=d, dilation_end=d, stride=stride)
    for seq_len, n in ((150, 3), (1000, 3), (40, 4), (16, 2)):
        for start, end in ((6, 1), (10, 3), (25, 1), (4, 4)):
            for total in (3, 5, 10, 20):
                for i_step in range(total):
                    for stride in (1, 3):
                        one(i_step=i_step, total=total, seq_len=seq_len, snippet_len=n, dilation_start=start,
                            dilation_end=end, stride=stride)
    one(i_step=0, total=1, seq_len=10, snippet_len=3, dilation_start=1, dilation_end=2, stride=1)     # gap_start < gap_end
    one(i_step=0, total=1, seq_len=10, snippet_len=3, dilation_start=0, dilation_end=0, stride=1)     # negative gap
    caps = [dict(seq_len=sl, snippet_len=n, dilation=d, capped=P.cap_max_dilation(seq_len=sl, snippet_len=n, dilation=d, verbose=False))
            for sl in (2, 3, 7, 10, 16, 31, 150, 1000) for n in (1, 2, 3, 4, 8) for d in (1, 2, 6, 10, 25, 400)]
    logging.disable(logging.NOTSET)
    with open(os.path.join(GOLD, "snippets_wide.json"), "w") as f:
        json.dump(dict(cases=cases, caps=caps), f, separators=(",", ":"))
    n_err = sum("error" in c for c in cases)
    print(f"snippets_wide: {len(cases)} scheduler cases ({n_err} of them invalid requests), {len(caps)} cap cases")


# --------------------------------------------------------------------------- co-alignment
def synth_snippets(N, H, W, w, dilations, seed, dtype=torch.float32, noise=0.01):
    """Aligner microbench generator of SURVEY.md section 8d (smooth base, per-snippet affine)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(N, 1, H // 4 + 2, W // 4 + 2, generator=g)
    base = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=False)
    base = 0.2 + 0.6 * (base - base.min()) / (base.max() - base.min())
    out = []
    for d in dilations:
        n = N - (w - 1) * d
        idx = torch.arange(n)[:, None] + torch.arange(w)[None, :] * d
        s = 1.0 + 0.3 * torch.rand(n, 1, 1, 1, 1, generator=g)
        t = 0.1 * torch.rand(n, 1, 1, 1, 1, generator=g)
        x = (base[idx] - t) / s + noise * torch.randn(n, w, 1, H, W, generator=g)
        out.append(x.to(dtype))
    return out


def gen_coalign(da, rp):
    from torch.optim.adam import Adam

    snaps = []

    class RecordingAdam(Adam):
        def step(self, closure=None):
            # state BEFORE the update + gradient at that state + params AFTER
            params = [p for g in self.param_groups for p in g["params"]]
            pre = [p.detach().clone() for p in params]
            st = [(self.state[p]["exp_avg"].clone(), self.state[p]["exp_avg_sq"].clone(),
                   int(self.state[p]["step"])) if len(self.state[p]) else None for p in params]
            loss = super().step(closure)
            snaps.append(dict(pre=pre, state=st, grad=[p.grad.detach().clone() for p in params],
                              post=[p.detach().clone() for p in params], loss=float(loss)))
            return loss

    da.Adam = RecordingAdam
    out = {}
    for tag, (N, H, W, w, dil, dtype, iters) in {
        "a": (14, 34, 44, 3, [1, 4], torch.float32, 40),
        "b": (20, 44, 54, 3, [1, 3, 6], torch.float32, 25),
        "h": (14, 34, 44, 3, [1, 4], torch.float16, 12),
        "w2": (9, 24, 34, 2, [1, 2], torch.float32, 10),
    }.items():
        snaps.clear()
        sn = synth_snippets(N, H, W, w, dil, seed=7, dtype=dtype)
        al = da.DepthAligner(device=torch.device("cpu"), num_iterations=iters)
        with contextlib.redirect_stdout(io.StringIO()):
            merged, scales, trans, hist = al.run(snippet_ls=[s.clone() for s in sn], dilations=list(dil))
        G = len(dil)
        out[f"{tag}_meta"] = np.array([N, H, W, w, iters, G] + list(dil), dtype=np.int64)
        for i, s in enumerate(sn):
            out[f"{tag}_snip{i}"] = s.numpy()
        out[f"{tag}_merged"] = merged.detach().numpy()
        for i in range(G):
            out[f"{tag}_scale{i}"] = scales[i].detach().reshape(-1).numpy()
            out[f"{tag}_trans{i}"] = trans[i].detach().reshape(-1).numpy()
        out[f"{tag}_hist"] = np.array(hist, dtype=np.float64)
        # teacher-forcing snapshots at a few iterations: params are [s_0..s_G-1, t_0..t_G-1]
        for it in sorted({0, 1, 5, iters - 1}):
            sp = snaps[it]
            for i in range(2 * G):
                out[f"{tag}_it{it}_pre{i}"] = sp["pre"][i].reshape(-1).numpy()
                out[f"{tag}_it{it}_grad{i}"] = sp["grad"][i].reshape(-1).numpy()
                out[f"{tag}_it{it}_post{i}"] = sp["post"][i].reshape(-1).numpy()
                if sp["state"][i] is not None:
                    out[f"{tag}_it{it}_m{i}"] = sp["state"][i][0].reshape(-1).numpy()
                    out[f"{tag}_it{it}_v{i}"] = sp["state"][i][1].reshape(-1).numpy()
            out[f"{tag}_it{it}_step"] = np.array(it, dtype=np.int64)
        print(f"coalign {tag}: loss {hist[0][0]:.6f} -> {hist[-1][0]:.6f}")
    np.savez_compressed(os.path.join(GOLD, "coalign.npz"), **out)


# --------------------------------------------------------------------------- attention processor
def gen_xattn(da, rp):
    from diffusers.models.attention_processor import Attention, AttnProcessor2_0

    out = {}
    torch.manual_seed(3)
    for tag, (C, heads, n, hw, cross) in {
        "self": (128, 2, 3, 40, None),
        "self_odd": (192, 3, 4, 21, None),
        "text": (128, 2, 3, 40, 96),
        "perframe": (128, 2, 3, 40, None),
    }.items():
        attn = Attention(query_dim=C, cross_attention_dim=cross, heads=heads, dim_head=64, bias=False)
        attn.set_processor(AttnProcessor2_0())
        x = torch.randn(n, hw, C)
        ctx = None if cross is None else torch.randn(1, 2, cross)
        nv = None if tag == "perframe" else n
        with torch.no_grad():
            y = attn(x, encoder_hidden_states=ctx, num_view=nv)
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_y"] = y.numpy()
        out[f"{tag}_meta"] = np.array([C, heads, n, hw, cross or 0, -1 if nv is None else nv], dtype=np.int64)
        if ctx is not None:
            out[f"{tag}_ctx"] = ctx.numpy()
        out[f"{tag}_wq"] = attn.to_q.weight.detach().numpy()
        out[f"{tag}_wk"] = attn.to_k.weight.detach().numpy()
        out[f"{tag}_wv"] = attn.to_v.weight.detach().numpy()
        out[f"{tag}_wo"] = attn.to_out[0].weight.detach().numpy()
        out[f"{tag}_bo"] = attn.to_out[0].bias.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "xattn.npz"), **out)
    print("xattn:", [k for k in out if k.endswith("_y")])


# --------------------------------------------------------------------------- DDIM
def gen_ddim(da, rp):
    from diffusers import DDIMScheduler

    out = {}
    g = torch.Generator().manual_seed(11)
    for tag, kw in {
        "vtrail": dict(prediction_type="v_prediction", timestep_spacing="trailing"),
        "epslead": dict(prediction_type="epsilon", timestep_spacing="leading", steps_offset=1),
    }.items():
        sch = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                            beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False, **kw)
        out[f"{tag}_alphas"] = sch.alphas_cumprod.numpy()
        for n in (1, 10, 20):
            sch.set_timesteps(n)
            ts = sch.timesteps.clone()
            out[f"{tag}_ts{n}"] = ts.numpy()
            for dt, dn in ((torch.float32, "f32"), (torch.float16, "f16")):
                x = torch.randn(3, 4, 6, 5, generator=g).to(dt)
                v = torch.randn(3, 4, 6, 5, generator=g).to(dt)
                for pos in sorted({0, len(ts) - 1}):
                    y = sch.step(v, ts[pos], x).prev_sample
                    out[f"{tag}_n{n}_{dn}_p{pos}_x"] = x.numpy()
                    out[f"{tag}_n{n}_{dn}_p{pos}_v"] = v.numpy()
                    out[f"{tag}_n{n}_{dn}_p{pos}_y"] = y.numpy()
                if n == 20:
                    z = sch.add_noise(x, v, ts[10])
                    out[f"{tag}_addnoise_{dn}_x"] = x.numpy()
                    out[f"{tag}_addnoise_{dn}_n"] = v.numpy()
                    out[f"{tag}_addnoise_{dn}_y"] = z.numpy()
                    out[f"{tag}_addnoise_t"] = np.array(int(ts[10]))
    np.savez_compressed(os.path.join(GOLD, "ddim.npz"), **out)
    print("ddim ok")


def gen_aligner_geometry(da, rp):
    """Index / shape bookkeeping of DepthAligner.run (depth_aligner.py:68-100) captured from the REFERENCE by intercepting its
    call of `optimize`: sequence length, cropped + subsampled sizes, triplet index tables -- for a grid of frame counts, window
    sizes, dilation sets, resolutions, borders and factors.  tests/golden/aligner_geometry.json; the CPU tier checks the
    product's `DepthAligner._problem` (rd_coalign_problem_init, host C code) and `create_triplet_indices` against it."""
    cases = []

    class Stop(Exception):
        pass

    def grab(self, snippet_ls, indices_list, sequence_length):
        raise Stop(dict(N=int(sequence_length), small=[list(t.shape) for t in snippet_ls],
                        idx=[snippet_digest(ix.tolist()) for ix in indices_list]))

    orig = da.DepthAligner.optimize
    da.DepthAligner.optimize = grab
    try:
        grid = []
        for (H, W) in ((44, 54), (47, 61), (30, 25), (256, 256)):
            for (border, factor) in ((2, 10), (1, 4), (3, 7), (2, 1)):
                for (N, w, dil) in ((7, 3, [1]), (12, 3, [1, 3]), (16, 3, [1, 4]), (16, 2, [1, 5, 8]), (21, 4, [1, 2, 5]),
                                    (40, 3, [1, 10, 19]), (9, 8, [1])):
                    grid.append((H, W, border, factor, N, w, dil))
        for (H, W) in ((432, 768), (576, 1024)):
            for (N, w, dil) in ((7, 3, [1, 2]), (9, 4, [1, 2])):
                grid.append((H, W, 2, 10, N, w, dil))
        for (H, W, border, factor, N, w, dil) in grid:
            snips = [torch.zeros(N - (w - 1) * d, w, 1, H, W) for d in dil]
            al = da.DepthAligner(device=torch.device("cpu"), factor=factor, border=border, num_iterations=1)
            try:
                al.run(snips, list(dil))
                raise RuntimeError("optimize was not reached")
            except Stop as st:
                cases.append(dict(H=H, W=W, border=border, factor=factor, w=w, dilations=dil, n=[len(t) for t in snips], **st.args[0]))
    finally:
        da.DepthAligner.optimize = orig
    with open(os.path.join(GOLD, "aligner_geometry.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print(f"aligner_geometry: {len(cases)} cases")


FORWARD_ARG_CASES = [
    # (seq_len, forward kwargs that differ from the defaults below)
    (16, dict(dilations=[2, 4])),
    (16, dict(dilations=[1, 4], snippet_lengths=[3, 3])),
    (16, dict(dilations=[1, 4, 6], snippet_lengths=[3, 4])),
    (16, dict(dilations=[1, 4, 6], init_infer_steps=[1, 2])),
    (16, dict(dilations=[1, 4], init_infer_steps=[0])),
    (16, dict(dilations=[1, 4], init_infer_steps=[1, 0])),
    (16, dict(dilations=[1, 4, 6], strides=[1, 2])),
    (16, dict(dilations=[1, 4], strides=[2])),
    (16, dict(dilations=[1, 4, 6], strides=[1, 2, 1])),
    (16, dict(dilations=[1, 10, 25])),
    (16, dict(dilations=[1, 10, 25], cap_dilation=False)),
    (150, dict(dilations=[1, 10, 25], refine_start_dilation=6)),
    (150, dict(dilations=[1, 25, 80], snippet_lengths=[3, 4, 2])),
    (10, dict(dilations=[1, 25], refine_start_dilation=6, refine_snippet_len=4)),
    (7, dict(dilations=[25, 1], snippet_lengths=[2, 3], init_infer_steps=[2, 1])),
]
FORWARD_ARG_DEFAULTS = dict(cap_dilation=True, snippet_lengths=[3], init_infer_steps=[1], strides=[1], coalign_kwargs=None,
                            refine_step=0, refine_snippet_len=3, refine_start_dilation=6, generator=None, verbose=False,
                            max_vae_bs=4, unload_snippet=False)


def gen_forward_args(da, rp):
    """Argument handling of RollingDepthPipeline.forward (rollingdepth_pipeline.py:213-258) recorded from the REFERENCE: the
    exception class and message of every rejected call, and for accepted calls the caller's `dilations` list after the in-place
    capping (the first stage after the checks, encode_rgb, is intercepted).  tests/golden/forward_args.json."""
    import copy
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import model_utils as mu
    from diffusers import AutoencoderKL, DDIMScheduler, UNet2DConditionModel
    unet = UNet2DConditionModel(**mu.TINY_UNET).eval()
    vae = AutoencoderKL(**mu.TINY_VAE).eval()
    pipe = rp.RollingDepthPipeline(unet=unet, vae=vae, scheduler=DDIMScheduler(**mu.DDIM_CFG), text_encoder=None, tokenizer=None)

    class Reached(Exception):
        pass

    def stop(*a, **k):
        raise Reached()
    pipe.encode_rgb = stop
    out = []
    for seq_len, kw in FORWARD_ARG_CASES:
        args = copy.deepcopy(dict(FORWARD_ARG_DEFAULTS, **kw))
        rec = dict(seq_len=seq_len, kwargs=copy.deepcopy(kw))
        try:
            pipe.forward(input_frames=torch.zeros(1, seq_len, 3, 16, 16), **args)
            raise RuntimeError("encode_rgb was not reached")
        except Reached:
            rec.update(accepted=True, dilations_after=list(args["dilations"]))
        except (AssertionError, NotImplementedError) as e:
            rec.update(accepted=False, error=type(e).__name__, message=str(e))
        out.append(rec)
    with open(os.path.join(GOLD, "forward_args.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(f"forward_args: {len(out)} calls, {sum(not r['accepted'] for r in out)} rejected")


DDIM_WIDE_CONFIGS = {
    # tag: constructor kwargs on top of num_train_timesteps=1000 (the first one is RollingDepth's scheduler_config.json)
    "rd": dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
               steps_offset=1, prediction_type="v_prediction", timestep_spacing="trailing"),
    "rd_one": dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=True,
                   steps_offset=0, prediction_type="v_prediction", timestep_spacing="trailing"),
    "sd_eps": dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
                   steps_offset=1, prediction_type="epsilon", timestep_spacing="leading"),
    "lin_eps_clip": dict(beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True, set_alpha_to_one=True,
                         steps_offset=0, prediction_type="epsilon", timestep_spacing="leading"),
    "lin_v_linspace": dict(beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=False, set_alpha_to_one=False,
                           steps_offset=0, prediction_type="v_prediction", timestep_spacing="linspace"),
    "sample_trail": dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                         set_alpha_to_one=False, steps_offset=1, prediction_type="sample", timestep_spacing="trailing"),
    "sample_linspace_clip": dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=True,
                                 clip_sample_range=0.8, set_alpha_to_one=True, steps_offset=0, prediction_type="sample",
                                 timestep_spacing="linspace"),
}
DDIM_WIDE_STEPS = (1, 2, 3, 7, 20, 50, 1000)


def gen_ddim_wide(da, rp):
    """The reference's DDIMScheduler over a grid of configurations (beta schedule, spacing, prediction type, clipping,
    set_alpha_to_one, steps_offset) x numbers of inference steps: timestep lists, and `step` outputs at the first / second /
    middle / last timestep on small fp32 and fp16 inputs.  tests/golden/ddim_wide.npz; the CPU tier checks the oracle's
    DDIMRef AND the product's host-side schedule (rollingdepth_b200/scheduler.py: timesteps + the four scalars it hands
    to rd_ddim_step) against it."""
    from diffusers import DDIMScheduler
    out = {}
    g = torch.Generator().manual_seed(23)
    x32 = torch.randn(2, 4, 3, 3, generator=g)
    v32 = torch.randn(2, 4, 3, 3, generator=g)
    out["x"], out["v"] = x32.numpy(), v32.numpy()
    n_steps = 0
    for tag, kw in DDIM_WIDE_CONFIGS.items():
        sch = DDIMScheduler(num_train_timesteps=1000, **kw)
        for n in DDIM_WIDE_STEPS:
            sch.set_timesteps(n)
            ts = sch.timesteps.clone()
            out[f"{tag}_ts{n}"] = ts.numpy()
            for pos in sorted({0, min(1, len(ts) - 1), len(ts) // 2, len(ts) - 1}):
                for dt, dn in ((torch.float32, "f32"), (torch.float16, "f16")):
                    try:
                        out[f"{tag}_n{n}_p{pos}_{dn}"] = sch.step(v32.to(dt), ts[pos], x32.to(dt)).prev_sample.numpy()
                        n_steps += 1
                    except IndexError:      # leading spacing + steps_offset with 1000 steps asks for timestep 1000
                        out[f"{tag}_n{n}_p{pos}_{dn}_indexerror"] = np.array(int(ts[pos]))
    np.savez_compressed(os.path.join(GOLD, "ddim_wide.npz"), **out)
    print(f"ddim_wide: {len(DDIM_WIDE_CONFIGS)} configurations x {len(DDIM_WIDE_STEPS)} step counts, {n_steps} step outputs")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="snippets,snippets_wide,forward_args,aligner_geometry,coalign,xattn,ddim,ddim_wide,pipeline")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    da, rp = import_reference()
    todo = args.only.split(",")
    if "snippets" in todo:
        gen_snippets(da, rp)
    if "snippets_wide" in todo:
        gen_snippets_wide(da, rp)
    if "forward_args" in todo:
        gen_forward_args(da, rp)
    if "aligner_geometry" in todo:
        gen_aligner_geometry(da, rp)
    if "coalign" in todo:
        gen_coalign(da, rp)
    if "xattn" in todo:
        gen_xattn(da, rp)
    if "ddim" in todo:
        gen_ddim(da, rp)
    if "ddim_wide" in todo:
        gen_ddim_wide(da, rp)
    if "pipeline" in todo:
        from oracle import gen_golden_pipeline
        gen_golden_pipeline.run(da, rp, GOLD)
    if "cfg1" in todo:        # not in the default list: minutes of CPU time (the full SD2 architecture)
        from oracle import gen_golden_cfg1
        gen_golden_cfg1.run(da, rp, GOLD)
    if "cfg1_loss_floor" in todo:   # appends the reference's own loss-curve sensitivity floor to the cfg1 fixture
        from oracle import gen_golden_cfg1
        gen_golden_cfg1.loss_floor(da, rp, GOLD)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
