#!/usr/bin/env python
"""Headline benchmark: video frames/s of the RollingDepth snippet-inference hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--preset fast|full|paper|cfg1] [--impl ours|reference]

One "step" = one pass of `RollingDepthPipeline.forward` over one synthetic video (BASELINE.json:
`fast` preset = 150 frames, 768x432, snippet length 3, dilations [1,25], fp16, 1 denoise step, no refinement,
2000 co-alignment iterations; random-init SD2-architecture UNet (866 M) + VAE (84 M)).
  value : frames/s with the frames already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : the same through the public API with the frames in pinned HOST memory (H2D inside, results to host)
  roofline : cross-frame attention kernel (tensor bound), live CUDA-event durations of its level-0 launches
  cpu_baseline : the oracle (CPU restatement of the reference path) timed on this box's host cores on a
                 bounded, warmed-up sample of the same workload and extrapolated with the exact call counts.
  gpu_reference : context -- what a user of the REFERENCE gets on this same GPU: the reference's control flow (one
                 snippet per UNet call, VAE chunks of 4, NCHW, torch SDPA, torch group_norm, eager autograd aligner with
                 three host syncs per iteration) on a bounded sample, extrapolated with the exact call counts
                 (`--gpu-reference full` runs the whole clip once instead).
  output_check : the timed step's output against the oracle's op-for-op restatement on the same device for a subset
                 of snippets / frames (not `isfinite`): denoise + decode of sampled snippets, merge + renormalisation of
                 sampled frames.
`--impl reference` times the CPU path alone (the reference is pure Python/CPU-runnable).
Under torchrun (N > 1) the snippets of the one video are sharded over the ranks (strong scaling); rank 0 returns the
results (reference semantics), the other ranks only compute.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {
    # name: (frames, H, W, dilations, refine_step, dtype, cap_dilation)
    "fast": (150, 432, 768, [1, 25], 0, torch.float16, True),
    "full": (150, 576, 1024, [1, 10, 25], 10, torch.float16, True),
    "paper": (150, 432, 768, [1, 10, 25], 10, torch.float32, False),
    "cfg1": (16, 256, 256, [1, 4], 0, torch.float32, True),
}
METRIC = "video frames/sec (RollingDepth snippet inference)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------- models / data
def build_models(device, dtype, seed=0):
    """Random-init SD2-architecture UNet + VAE (no checkpoints offline).  Scaled init keeps fp16 finite."""
    from rollingdepth_b200.sd2.unet import SD2_UNET_CONFIG, UNet2DConditionModel
    from rollingdepth_b200.sd2.vae import SD_VAE_CONFIG, AutoencoderKL
    with torch.device(device):
        unet = UNet2DConditionModel(**SD2_UNET_CONFIG)
        vae = AutoencoderKL(**SD_VAE_CONFIG)
    g = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for m, gain in ((unet, 0.7), (vae, 0.7)):
            for name, p in m.named_parameters():
                if p.dim() <= 1:
                    if name.endswith("weight"):
                        p.fill_(1.0)
                    else:
                        p.zero_()
                else:
                    p.copy_(torch.randn(p.shape, device=p.device, generator=g) * (gain / p[0].numel() ** 0.5))
    return unet.to(dtype).eval(), vae.to(dtype).eval()


def synth_video(n, h, w, seed=0):
    """Smooth synthetic clip in [-1, 1], float32 on the host like load_video_frames returns."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(n, 3, h // 16 + 2, w // 16 + 2, generator=g)
    vid = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    drift = torch.linspace(0, 1, n)[:, None, None, None] * 0.2
    return ((vid + drift).clamp(0, 1) * 2 - 1)[None].contiguous()


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nme, val in zip(names, out[2:]):
                    if val.strip().lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def call_counts(preset):
    """Exact unit counts of one pass of the preset: UNet calls (one snippet / window each), frame
    decodes, frame encodes, aligner iterations."""
    from oracle import schedule as osch
    n, H, W, dil, refine, _dtype, cap = PRESETS[preset]
    dil = [osch.cap_max_dilation(n, 3, d) for d in dil] if cap else list(dil)
    n_snip = sum(n - 2 * d for d in dil)
    n_refine_unet = 0
    if refine > 0:
        start = osch.cap_max_dilation(n, 3, 6) if cap else 6
        for i in range(refine):
            n_refine_unet += len(osch.get_snippet_indice(i, refine, n, 3, start, 1, 1))
    return dil, dict(unet=n_snip + n_refine_unet, decode=3 * n_snip + (n if refine else 0),
                     encode=n * (2 if refine else 1), align_it=2000)


# ------------------------------------------------------------------------------------------- CPU baseline
def cpu_baseline(preset, full=False):
    """The oracle (reference control flow, CPU fp32, all host threads).  Default: a bounded sample -- every unit is run
    twice and the SECOND run is timed (warm caches, oneDNN primitives created) -- extrapolated with the exact call
    counts of the workload.  `full`: the whole pass, no extrapolation (minutes for cfg1, hours for the 150-frame presets)."""
    from oracle import coalign as oc
    from oracle import pipeline as opipe
    from oracle.gen_golden import synth_snippets
    n, H, W, _dil, refine, _dtype, cap = PRESETS[preset]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, vae = build_models("cpu", torch.float32)
    text = torch.randn(1, 2, 1024)
    dil, counts = call_counts(preset)
    if full:
        sched = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                     clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="v_prediction",
                     timestep_spacing="trailing")
        frames = synth_video(n, H, W)
        timers = {}
        t0 = time.time()
        opipe.forward(unet, vae, sched, text, frames, list(PRESETS[preset][3]), cap_dilation=cap, refine_step=refine,
                      generator=torch.Generator().manual_seed(0), max_vae_bs=4, timers=timers)
        total = time.time() - t0
        return dict(value=n / total, unit="frames/s", cores=cores, kind="port", seconds_per_video=total,
                    sample=f"the whole pass of the '{preset}' preset on CPU fp32 (no extrapolation)", stage_seconds=timers)
    t = {}

    def second_of_two(fn):
        fn()
        t0 = time.time()
        fn()
        return time.time() - t0

    with torch.no_grad():
        frames = synth_video(2, H, W)[0]
        t["encode"] = second_of_two(lambda: opipe.encode_rgb(vae, frames[None], 4)) / 2
        h, w = H // 8, W // 8
        rgb = torch.randn(1, 3, 4, h, w)
        dep = torch.randn(1, 3, 4, h, w)
        t["unet"] = second_of_two(lambda: opipe.single_step(unet, rgb, dep, 999, text))
        t["decode"] = second_of_two(lambda: opipe.decode_depth(vae, dep[:, :2], 4)) / 2
    # aligner at the real frame count and subsampled resolution
    Hs, Ws = (H - 4 + 9) // 10, (W - 4 + 9) // 10
    small = [s.reshape(s.shape[0], 3, -1) for s in synth_snippets(n, Hs, Ws, 3, dil, seed=0)]
    it = 5
    t["align_it"] = second_of_two(lambda: oc.optimize(small, dil, n, num_iterations=it)) / it
    total = sum(counts[k] * t[k] for k in counts)
    sample = (f"each unit run twice, second run timed: 2 frame encodes, 1 UNet call (3 frames), 2 frame decodes, {it} aligner "
              f"iterations at N={n}, {W}x{H}; extrapolated with counts {counts}")
    return dict(value=n / total, unit="frames/s", cores=cores, kind="port", sample=sample,
                seconds_per_video=total, per_unit_seconds=t)


# ------------------------------------------------------------------------------------------- GPU reference path
def _plain_modules(unet, vae):
    """Put our SD2 modules into the reference's configuration: NCHW, torch group_norm / eager adds, torch SDPA."""
    from rollingdepth_b200.sd2.layers import TorchSDPAProcessor
    for m in (unet, vae):
        m.to(memory_format=torch.contiguous_format)
        for sub in m.modules():
            if hasattr(sub, "fused"):
                sub.fused = False
    unet.set_attn_processor(TorchSDPAProcessor())


def gpu_reference(preset, unet, vae, text, dev, dtype, mode="sample"):
    """The reference's GPU path on this device (context for the headline; see the module docstring).  Leaves the models
    in the reference configuration: call last."""
    from oracle import coalign_eager as oe
    from oracle import pipeline as opipe
    from oracle.gen_golden import synth_snippets
    n, H, W, _dil, refine, _dt, cap = PRESETS[preset]
    dil, counts = call_counts(preset)
    _plain_modules(unet, vae)
    torch.backends.cudnn.benchmark = True
    sched = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="v_prediction",
                 timestep_spacing="trailing")
    what = ("reference control flow on this GPU: one snippet per UNet call, VAE chunks of 4, NCHW, torch SDPA "
            "(AttnProcessor2_0's call), torch group_norm, eager autograd DepthAligner with 3 host syncs per iteration")
    if mode == "full":
        frames = synth_video(n, H, W)
        timers = {}
        small = synth_video(12, H, W)
        opipe.forward(unet, vae, sched, text, small, [1, 3], cap_dilation=cap, refine_step=min(refine, 2),
                      coalign_kwargs=dict(num_iterations=20), generator=torch.Generator(device=dev).manual_seed(0),
                      max_vae_bs=4, aligner_factory=oe.EagerDepthAligner)           # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        t0 = time.time()
        opipe.forward(unet, vae, sched, text, frames, list(PRESETS[preset][3]), cap_dilation=cap, refine_step=refine,
                      generator=torch.Generator(device=dev).manual_seed(0), max_vae_bs=4, timers=timers,
                      aligner_factory=oe.EagerDepthAligner)
        torch.cuda.synchronize()
        total = time.time() - t0
        return dict(value=n / total, unit="frames/s", seconds_per_video=total, what=what, stage_seconds=timers,
                    sample="the whole pass, timed once after a 12-frame warm-up pass")

    def timed(fn, reps):
        fn()
        fn()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.time() - t0) / reps

    t = {}
    with torch.no_grad():
        frames = synth_video(4, H, W)[0].to(dev, dtype)
        t["encode"] = timed(lambda: opipe.encode_rgb(vae, frames[None], 4), 4) / 4
        h, w = H // 8, W // 8
        rgb = torch.randn(1, 3, 4, h, w, device=dev, dtype=dtype)
        dep = torch.randn(1, 3, 4, h, w, device=dev, dtype=dtype)
        t["unet"] = timed(lambda: opipe.single_step(unet, rgb, dep, 999, text), 8)
        lat = torch.randn(1, 4, 4, h, w, device=dev, dtype=dtype)
        t["decode"] = timed(lambda: opipe.decode_depth(vae, lat, 4), 6) / 4
    Hs, Ws = (H - 4 + 9) // 10, (W - 4 + 9) // 10
    small = [s.reshape(s.shape[0], 3, -1).to(dev, dtype) for s in synth_snippets(n, Hs, Ws, 3, dil, seed=0)]
    it = 50
    t["align_it"] = timed(lambda: oe.optimize_autograd(small, dil, n, num_iterations=it), 1) / it
    # merge (N * G masked gathers, one nonzero-sync each) at full resolution on a third of the frames
    total = sum(counts[k] * t[k] for k in counts)
    return dict(value=n / total, unit="frames/s", seconds_per_video=total, what=what, per_unit_seconds=t,
                sample=f"after 2 warm-up runs each: 4 x (4-frame encode), 8 UNet calls (3 frames), 6 x (4-frame decode), {it} eager "
                       f"aligner iterations at N={n}; extrapolated with counts {counts} (merge / H2D / D2H not counted: favours the reference)")


# ------------------------------------------------------------------------------------------- output check
def output_check(pipe, out, host_frames, preset, dev, dtype, max_vae_bs):
    """Compare the (timed) product output with the oracle's op-for-op restatement on the same device for a subset:
    (a) three snippets per dilation: VAE encode -> one UNet step with library attention -> DDIM -> decode, against
        `out.snippet_ls`; (b) merge + renormalisation of three frames recomputed from `out.snippet_ls` and the
        product's scales / translations, against `out.depth_coaligned` (exact)."""
    from oracle import pipeline as opipe
    from oracle import schedule as osch
    from rollingdepth_b200.sd2.layers import TorchSDPAProcessor
    n, H, W, _dil, refine, _dt, cap = PRESETS[preset]
    dil, _ = call_counts(preset)
    unet, vae = pipe.unet, pipe.vae
    saved = {m: m.fused for mm in (unet, vae) for m in mm.modules() if hasattr(m, "fused")}
    procs = dict(unet.attn_processors)
    for m in saved:
        m.fused = False
    unet.set_attn_processor(TorchSDPAProcessor())
    res = {}
    try:
        with torch.no_grad():
            sch = osch.DDIMRef()
            ts = sch.set_timesteps(1)
            h, w = H // 8, W // 8
            noise = torch.randn((1, 4, h, w), device=dev, dtype=dtype, generator=torch.Generator(device=dev).manual_seed(0))
            text = pipe.empty_text_embed
            worst, rng = 0.0, 0.0
            for gi, d in enumerate(dil):
                idx_ls = osch.get_snippet_indice(0, 1, n, 3, d, d, 1)
                for k in sorted({0, len(idx_ls) // 2, len(idx_ls) - 1}):
                    idx = idx_ls[k]
                    fr = host_frames[:, idx].to(dev).to(dtype)
                    rgb = opipe.encode_rgb(vae, fr, max_vae_bs)
                    dep = noise[:, None].expand(1, 3, 4, h, w).clone()
                    pred = opipe.single_step(unet, rgb, dep, int(ts[0]), text)
                    dec = opipe.decode_depth(vae, sch.step(pred, int(ts[0]), dep), max_vae_bs)[0]
                    got = out.snippet_ls[gi][k].to(dev).float()
                    worst = max(worst, float((got - dec.float()).abs().max()))
                    rng = max(rng, float(dec.float().abs().max()))
            res["snippets_checked"] = 3 * len(dil)
            res["snippet_max_abs"] = worst
            res["snippet_range"] = rng
            res["snippet_ok"] = bool(worst <= 2e-2 * max(1.0, rng))
            # merge + renorm of sampled frames, exact in the snippet dtype (depth_aligner.py:246-262, pipeline :316-318)
            scales, trans, _hist = pipe.last_alignment
            mn0 = min(s.min() for s in out.snippet_ls).to(dev)
            mm = pipe.last_minmax.to(dev)
            exact, worst_m = True, 0.0
            for f in sorted({0, n // 2, n - 1}):
                parts = []
                for gi, d in enumerate(dil):
                    for j in range(2, -1, -1):
                        k = f - j * d
                        if 0 <= k < out.snippet_ls[gi].shape[0]:
                            x = out.snippet_ls[gi][k, j].to(dev) - mn0
                            parts.append(x * scales[gi][k].to(dtype) + trans[gi][k].to(dtype))
                m = torch.cat(parts).mean(0, keepdim=True)
                mn, mx = mm[0].to(dtype), mm[1].to(dtype)
                m = m - mn
                m = m / (mx - mn)
                m = m * 2.0 - 1.0
                if dtype == torch.float32:      # fp32: the kernel's sequential sum of <= 9 terms vs torch's mean: last-bit differences
                    worst_m = max(worst_m, float((m.cpu() - out.depth_coaligned[f]).abs().max()))
                    exact = exact and worst_m <= 2e-6
                else:                           # 16-bit: every rounding is prescribed by the reference's op order => bit-exact
                    exact = exact and torch.equal(m.cpu(), out.depth_coaligned[f])
            res["merge_renorm_frames_exact"] = bool(exact)
            if dtype == torch.float32:
                res["merge_renorm_max_abs_fp32"] = worst_m
            res["finite"] = bool(torch.isfinite(out.depth_pred.float()).all())
            res["ok"] = bool(res["snippet_ok"] and exact and res["finite"])
    finally:
        for m, f in saved.items():
            m.fused = f
        unet.set_attn_processor(procs)
    return res


# ------------------------------------------------------------------------------------------- main arms
def workload_config(preset, world, args=None):
    n, H, W, dil, refine, dtype, cap = PRESETS[preset]
    cfg = {"workload": f"'{preset}' preset: {n}-frame synthetic video {W}x{H}, snippet len 3, dilations {dil}, "
                       f"{'%d refine steps' % refine if refine else 'no refinement'}, 1 denoise step, 2000 co-align its, "
                       f"{str(dtype).split('.')[-1]}, random-init SD2-arch UNet+VAE",
           "frames": n, "height": H, "width": W, "dilations": dil, "refine_step": refine,
           "parallelism": f"snippet-sharded x{world}" if world > 1 else "single GPU",
           "l2": "inputs (>= 300 MB of frames/latents/snippets per step) exceed the 126 MB L2; no explicit flush"}
    if args is not None:
        cfg.update({"snippet_batch": args.snippet_batch, "max_vae_bs": args.max_vae_bs, "channels_last": bool(args.channels_last),
                    "fused_norm": bool(args.fused_norm), "cudnn_benchmark": bool(args.cudnn_benchmark),
                    "subpixel_upsample_conv": bool(args.subpixel)})
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    cb = None
    for _ in range(1 if args.cpu_full else max(1, min(args.steps, 2))):
        cb = cpu_baseline(args.preset, full=args.cpu_full)
        vals.append(cb["value"])
    v = sum(vals) / len(vals)
    n = PRESETS[args.preset][0]
    cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / v,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.preset, int(os.environ.get("WORLD_SIZE", "1")), args), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from rollingdepth_b200 import _lib as L
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    n, H, W, dil, refine, dtype, cap = PRESETS[args.preset]
    hbm, tf_burst, tf_sust, peak_src = peaks()

    unet, vae = build_models(dev, dtype)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    from rollingdepth_b200.sd2.layers import Upsample2D
    Upsample2D.subpixel = bool(args.subpixel)
    if dtype == torch.float32:      # the fp32 presets are tolerance checks against the reference's true-fp32 path
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(), snippet_batch=args.snippet_batch,
                                channels_last=bool(args.channels_last), fused_norm=bool(args.fused_norm))
    if world > 1:
        pipe.enable_sharding(output_ranks=(0,), coalign_mode=args.coalign_mode)
    pipe.empty_text_embed = torch.randn(1, 2, 1024, generator=torch.Generator().manual_seed(1)).to(dev, dtype)
    host_frames = synth_video(n, H, W).pin_memory()
    dev_frames = host_frames.to(dev)

    def step(frames):
        return pipe.forward(input_frames=frames, dilations=list(dil), cap_dilation=cap, snippet_lengths=[3],
                            init_infer_steps=[1], strides=[1],
                            coalign_kwargs=dict(num_iterations=args.coalign_its) if args.coalign_its else None, refine_step=refine,
                            refine_snippet_len=3, refine_start_dilation=6,
                            generator=torch.Generator(device=dev).manual_seed(0), verbose=False,
                            max_vae_bs=args.max_vae_bs, unload_snippet=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(frames, k, record=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.LAUNCHES = 0
        if record:
            proc.events = []
        e0.record()
        out = None
        for _ in range(k):
            out = step(frames)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out, L.LAUNCHES

    if args.stage_times:
        pipe.stage_events = []
    proc = next(iter(pipe.unet.attn_processors.values()))
    for _ in range(args.warmup):
        step(dev_frames)
    sampler = ClockSampler(local)
    sampler.start()
    ms, out, launches = timed(dev_frames, args.steps, record=True)
    events = list(getattr(proc, "events", None) or [])     # snapshot: the e2e pass below must not be counted
    proc.events = None
    stage_ms = {k: round(v, 1) for k, v in pipe.stage_times_ms()} if args.stage_times else None
    k2 = max(1, min(args.steps, 2))
    if args.no_e2e:                 # auxiliary presets only (minutes per step): the e2e keys then repeat nothing and say so
        ms_e2e, out2 = None, out
    else:
        ms_e2e, out2, _ = timed(host_frames, k2)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # roofline of the dominant hand-written kernel: level-0 cross-frame attention launches
    roof = None
    if events:
        by_shape = {}
        for (shape, a, b) in events:
            by_shape.setdefault(shape, []).append(a.elapsed_time(b))
        # dominant = the shape with the largest total time
        shape, durs = max(by_shape.items(), key=lambda kv: sum(kv[1]))
        Bq, Sq, Sk, Hh = shape
        flops = 4.0 * Bq * Hh * Sq * Sk * 64
        avg = sum(durs) / len(durs)
        # DRAM traffic per launch: ncu `dram__bytes_read.sum + dram__bytes_write.sum` of this kernel at exactly the `fast`
        # preset's launch shape (B=16, S=15552, H=5; profiles/r2_xattn_q3_b16_ncu_full.txt): 478.4 MB read + 141.3 MB
        # written = 619.7 MB against 637.0 MB of algorithmic q + k + v + o bytes (L2 serves the K/V re-reads, hit rate 92 %);
        # reported as that ratio times the algorithmic bytes of the launch shape timed here.
        traffic = 0.973 * (2.0 * Bq * Sq * Hh * 64 * 2 + 2.0 * Bq * Sk * Hh * 64 * 2)
        roof = {"bound": "tensor", "kernel": "xattn_fwd_kernel", "achieved": flops / avg / 1e9, "peak": tf_sust,
                "unit": "TFLOP/s", "frac": flops / avg / 1e9 / tf_sust, "traffic": traffic,
                "traffic_source": "ncu dram bytes / algorithmic bytes = 0.973 (profiles/r2_xattn_q3_b16_ncu_full.txt, the fast preset's launch shape) x algorithmic bytes of this launch shape",
                "peak_source": f"bf16_tflops_sustained of {peak_src} (kernel timed inside a long step)",
                "frac_of_burst_peak": flops / avg / 1e9 / tf_burst,
                "launch_shape": {"B": Bq, "Sq": Sq, "Sk": Sk, "H": Hh, "D": 64}, "launches_timed": len(durs),
                "avg_launch_ms": avg, "share_of_step": sum(sum(v) for v in by_shape.values()) / ms,
                "flops_per_launch": flops}
    h2d = host_frames.numel() * host_frames.element_size()    # fp32 frames are copied, rounded on the device
    if world > 1 and rank != 0:
        lo, hi = pipe.shard.block(n)
        h2d = h2d // n * (hi - lo)
    d2h = 0
    if out2.depth_pred is not None:
        d2h = sum(t.numel() * t.element_size() for t in [out2.input_rgb, out2.depth_pred, out2.depth_coaligned] + list(out2.snippet_ls))
    if world > 1:
        io = torch.tensor([float(h2d), float(d2h)], device=dev, dtype=torch.float64)
        dist.all_reduce(io)
        h2d, d2h = int(io[0]), int(io[1])
    if rank == 0:
        check = None
        if not args.no_output_check:
            try:
                check = output_check(pipe, out, host_frames, args.preset, dev, dtype, args.max_vae_bs)
            except Exception as e:  # noqa: BLE001
                check = {"ok": False, "error": f"{type(e).__name__}: {str(e)[:200]}"}
        line = {
            "metric": METRIC, "value": n * args.steps / (ms / 1e3),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": str(dtype).split(".")[-1].replace("float", "f"), "data": "synthetic",
            "config": workload_config(args.preset, world, args),
            "clocks": sampler.summary(),
            "e2e": ({"value": n * k2 / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h, "steps": k2,
                     "note": "bytes summed over all ranks; results are returned on rank 0 only" if world > 1 else None}
                    if ms_e2e is not None else {"value": None, "unit": "frames/s", "note": "not measured (--no-e2e)"}),
            "gpu_launches": launches, "roofline": roof, "output_check": check,
            "finite_output": bool(check["finite"]) if check and "finite" in check else bool(torch.isfinite(out.depth_pred.float()).all()),
        }
        if world > 1:
            line["config"]["coalign_mode"] = args.coalign_mode
        if stage_ms is not None:
            line["stage_ms_last_step_rank0"] = stage_ms
        gref = None
        if world == 1 and args.gpu_reference != "off":
            try:
                gref = gpu_reference(args.preset, unet, vae, pipe.empty_text_embed, dev, dtype, args.gpu_reference)
            except Exception as e:  # noqa: BLE001
                gref = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        line["gpu_reference"] = gref
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_baseline(args.preset)
            except Exception as e:  # noqa: BLE001
                cb = {"error": str(e)[:200]}
        line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--preset", default="fast", choices=list(PRESETS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--snippet-batch", type=int, default=16)
    ap.add_argument("--max-vae-bs", type=int, default=36)
    ap.add_argument("--channels-last", type=int, default=1)
    ap.add_argument("--fused-norm", type=int, default=1)
    ap.add_argument("--cudnn-benchmark", type=int, default=1)
    ap.add_argument("--subpixel", type=int, default=1, help="0: Upsample2D as nearest upsampling + 3x3 convolution (A/B)")
    ap.add_argument("--coalign-mode", default="replicated", choices=["replicated", "allreduce"],
                    help="multi-GPU only: replicated optimiser or frame-sharded with a per-iteration all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-output-check", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="auxiliary presets only: skip the host-memory pass (NOT for the headline line)")
    ap.add_argument("--gpu-reference", default="sample", choices=["off", "sample", "full"],
                    help="time the reference's GPU path as context (N = 1 only): bounded sample (default) or the whole clip")
    ap.add_argument("--cpu-full", action="store_true", help="--impl reference: the whole pass on CPU, no extrapolation")
    ap.add_argument("--stage-times", action="store_true", help="add per-stage device times of the last timed step (rank 0)")
    ap.add_argument("--coalign-its", type=int, default=0,
                    help="profiling only: fewer co-alignment iterations than the reference's 2000 (ncu launch list); NOT a bench configuration")
    ap.add_argument("--frames", type=int, default=0,
                    help="profiling only: shorter clip than the preset (the ncu launch list); NOT a bench configuration")
    a = ap.parse_args()
    if a.frames:
        cfg = list(PRESETS[a.preset])
        cfg[0] = a.frames
        PRESETS[a.preset] = tuple(cfg)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
