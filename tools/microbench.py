"""Kernel microbenchmarks (BASELINE.json config 5): cross-frame attention sweep and the 1000-frame
co-alignment problem.  Prints one JSON object per line; CUDA-event timing after warm-up.

    python tools/microbench.py attn [--full]      python tools/microbench.py align [--iters 200]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def time_cuda(fn, warmup=3, iters=10, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_attn(args):
    """Cross-frame attention: the preset shapes at B = 1 and at the snippet batch the pipeline runs (B = 16 / 4), and with
    --full the config-5 sweep (n = 3..8 frames x 512 / 768 / 1024 px, levels 0 and 1, B = 1).  Comparator on the same
    strided views: torch SDPA (the reference's call) -- default backend choice and, explicitly, cuDNN and flash."""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    from rollingdepth_b200 import ops
    dev = torch.device("cuda:0")
    hbm, tf_peak, src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cases = []
    if args.full:
        for res in (512, 768, 1024):
            for n in (3, 4, 5, 6, 7, 8):
                hw = (res // 8) ** 2
                cases.append((f"{res}px n={n} L0", 1, n * hw, 5))
                cases.append((f"{res}px n={n} L1", 1, n * hw // 4, 10))
    cases += [("fast L0 (768x432 n=3)", 1, 15552, 5), ("fast L1", 1, 3888, 10), ("fast L2", 1, 1008, 20), ("fast mid", 1, 252, 20),
              ("full L0 (1024x576 n=3)", 1, 27648, 5), ("full L1", 1, 6912, 10),
              ("fast L0 B=16", 16, 15552, 5), ("fast L1 B=16", 16, 3888, 10), ("fast L2 B=16", 16, 1008, 20),
              ("full L0 B=16", 16, 27648, 5), ("full L1 B=16", 16, 6912, 10), ("full L0 B=4", 4, 27648, 5)]
    if args.only:
        cases = [c for c in cases if args.only in c[0]]
    backends = [("torch_sdpa", None), ("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)]
    for dtype in ((torch.float16,) if (args.only or args.full) else (torch.float16, torch.bfloat16)):
        for name, B, S, H in cases:
            q, k, v = [torch.randn(B, S, H * 64, device=dev, dtype=dtype).view(B, S, H, 64) for _ in range(3)]
            out = torch.empty_like(q)
            flops = 4.0 * B * H * S * S * 64
            iters = 10 if flops < 3e13 else 5
            for flags in args.flags:
                med, best = time_cuda(lambda: ops.xattn_fwd(q, k, v, out=out, flags=flags), iters=iters, flush=flush)
                rec = dict(bench="xattn", case=name, B=B, S=S, H=H, dtype=str(dtype), flags=flags, ms=med, ms_best=best,
                           tflops=flops / med / 1e9, frac_of_peak=flops / med / 1e9 / tf_peak, peak=tf_peak, peak_src=src)
                if flags == args.flags[0] and not args.no_sdpa:
                    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
                    for bname, be in backends:
                        try:
                            if be is None:
                                m2, _ = time_cuda(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt),
                                                  iters=iters, flush=flush)
                            else:
                                with sdpa_kernel(be):
                                    m2, _ = time_cuda(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt),
                                                      iters=iters, flush=flush)
                            rec[f"{bname}_ms"] = m2
                            rec[f"{bname}_tflops"] = flops / m2 / 1e9
                        except Exception as e:  # noqa: BLE001
                            rec[f"{bname}_err"] = str(e)[:80]
                    best_lib = max([rec.get(f"{b}_tflops", 0.0) for b, _ in backends])
                    rec["vs_best_library"] = rec["tflops"] / best_lib if best_lib > 0 else None
                print(json.dumps(rec), flush=True)


def bench_attn_fp32(args):
    """The fp32 attention path (`paper` preset / cfg1: a plain FFMA kernel, <= 1e-4) against torch SDPA in fp32."""
    from rollingdepth_b200 import ops
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    for name, B, S, H in [("cfg1 L0 (256x256 n=3)", 1, 3072, 5), ("paper L0 (768x432 n=3)", 1, 15552, 5), ("paper L1", 1, 3888, 10)]:
        q, k, v = [torch.randn(B, S, H * 64, device=dev).view(B, S, H, 64) for _ in range(3)]
        out = torch.empty_like(q)
        flops = 4.0 * B * H * S * S * 64
        from rollingdepth_b200 import _lib as L
        med, _ = time_cuda(lambda: ops.xattn_fwd(q, k, v, out=out, flags=L.XATTN_F32(1)), iters=5)
        ref = out.clone()
        med3, _ = time_cuda(lambda: ops.xattn_fwd(q, k, v, out=out, flags=L.XATTN_F32(2)), iters=5)
        diff = float((out - ref).abs().max())
        qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        m2, _ = time_cuda(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), iters=5)
        print(json.dumps(dict(bench="xattn_fp32", case=name, S=S, H=H, ms=med, tflops=flops / med / 1e9, ms_3xtf32=med3,
                              tflops_3xtf32=flops / med3 / 1e9, max_abs_3xtf32_vs_ffma=diff, torch_sdpa_fp32_ms=m2,
                              torch_sdpa_fp32_tflops=flops / m2 / 1e9)), flush=True)


def bench_align(args):
    from rollingdepth_b200.depth_aligner import DepthAligner
    from oracle.gen_golden import synth_snippets  # generator only (inputs), not a compute path
    dev = torch.device("cuda:0")
    hbm, tf_peak, src = peaks()
    cfgs = [("1000f 768x432 [1,25] fp32", 1000, 432, 768, [1, 25], torch.float32),
            ("1000f 768x432 [1,25] fp16", 1000, 432, 768, [1, 25], torch.float16),
            ("1000f 1024x576 [1,10,25] fp32", 1000, 576, 1024, [1, 10, 25], torch.float32),
            ("150f 768x432 [1,25] fp16", 150, 432, 768, [1, 25], torch.float16)]
    import ctypes as C
    from rollingdepth_b200 import _lib as L
    for name, N, H, W, dil, dtype in cfgs:
        if args.small and N > 200:
            continue
        if args.only and args.only not in name:
            continue
        al = DepthAligner(device=dev, num_iterations=args.iters)
        # Build the subsampled problem directly on the device (full-resolution 1000-frame snippets would be
        # 15-20 GB); the optimiser only ever sees x_small.  prep / min / merge are timed separately below.
        pb = L.CoalignProblem()
        pb.G, pb.w, pb.N, pb.H, pb.W, pb.border, pb.factor = len(dil), 3, N, H, W, 2, 10
        for g_, d_ in enumerate(dil):
            pb.d[g_] = d_
        pb.dtype = L.dtype_code(dtype)
        L.check(L.lib().rd_coalign_problem_init(C.byref(pb)), "init")
        sn_small = synth_snippets(N, pb.Hs, pb.Ws, 3, dil, seed=0, dtype=dtype)
        mn = min(float(s.min()) for s in sn_small)
        x_small = torch.zeros(pb.n_total, 3, pb.P_pad, dtype=dtype, device=dev)
        x_small[:, :, :pb.P] = torch.cat([(s - mn).reshape(s.shape[0], 3, -1) for s in sn_small]).to(dev)
        state = al.new_state(pb)
        al.optimize_state(pb, x_small, state, 0, 5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hist = al.optimize_state(pb, x_small, state, 5, args.iters)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        es = 4 if dtype == torch.float32 else 2
        bytes_iter = pb.n_total * pb.w * pb.P * es
        gbs = bytes_iter * args.iters / ms / 1e6
        h = hist.cpu()
        print(json.dumps(dict(bench="coalign", case=name, n_snip=pb.n_total, P=pb.P, iters=args.iters, ms_total=ms,
                              us_per_iter=1e3 * ms / args.iters, MB_per_iter=bytes_iter / 1e6, GBs=gbs,
                              frac_of_hbm=gbs / hbm, peak=hbm, peak_src=src, loss0=float(h[0, 0]), lossN=float(h[-1, 0]),
                              fits_L2=bytes_iter < 126e6)), flush=True)
    if args.only:
        return
    # merge / min / prep at full resolution (150 frames, fp16, fast preset)
    N, H, W, dil, dtype = 150, 432, 768, [1, 25], torch.float16
    al = DepthAligner(device=dev)
    snips = [torch.rand(N - 2 * d, 3, 1, H, W, device=dev, dtype=dtype) for d in dil]
    tot_bytes = sum(s.numel() for s in snips) * 2
    pb, dsn, dev_min, x_small = al.prepare(snips, dil)
    state = al.new_state(pb)
    med, best = time_cuda(lambda: al.prepare(snips, dil))
    print(json.dumps(dict(bench="min+prep", case="150f 768x432 fp16", ms=med, GBs=tot_bytes / med / 1e6, frac_of_hbm=tot_bytes / med / 1e6 / hbm)))
    med, best = time_cuda(lambda: al.merge(pb, dsn, dev_min, state[0], state[1]))
    mb = tot_bytes + N * H * W * 2
    print(json.dumps(dict(bench="merge", case="150f 768x432 fp16", ms=med, GBs=mb / med / 1e6, frac_of_hbm=mb / med / 1e6 / hbm)))


def bench_gn(args):
    """Channels-last element-wise kernels on VAE-decoder-sized activations (12 frames of the fast preset)."""
    from rollingdepth_b200 import ops
    dev = torch.device("cuda:0")
    hbm, _, src = peaks()
    cases = [("dec L3 12x432x768x128", 12, 432, 768, 128), ("dec L2 12x216x384x256", 12, 216, 384, 256),
             ("dec L1 12x108x192x512", 12, 108, 192, 512)]
    if args.full:       # UNet activations of one 16-snippet batch (48 frames) and the small decoder levels
        cases += [("unet L0 48x54x96x320", 48, 54, 96, 320), ("unet L0 cat 48x54x96x640", 48, 54, 96, 640),
                  ("unet L1 48x27x48x640", 48, 27, 48, 640), ("unet L1 cat 48x27x48x1280", 48, 27, 48, 1280),
                  ("unet L2 48x14x24x1280", 48, 14, 24, 1280), ("unet L3 cat 48x7x12x2560", 48, 7, 12, 2560),
                  ("dec L0 12x54x96x512", 12, 54, 96, 512)]
    for name, N, H, W, Cc in cases:
        x = torch.randn(N, Cc, H, W, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        res = torch.randn_like(x)
        w = torch.ones(Cc, device=dev, dtype=torch.float16)
        b = torch.zeros(Cc, device=dev, dtype=torch.float16)
        nbytes = x.numel() * 2
        if args.only and args.only not in name:
            continue
        med, best = time_cuda(lambda: ops.group_norm_nhwc(x, 32, w, b, 1e-6, True))
        print(json.dumps(dict(bench="groupnorm+silu", case=name, ms=med, algorithmic_GB=3 * nbytes / 1e9,
                              GBs=3 * nbytes / med / 1e6, frac_of_hbm=3 * nbytes / med / 1e6 / hbm, peak_src=src)), flush=True)
        med, best = time_cuda(lambda: ops.group_norm_nhwc(x, 32, w, b, 1e-6, False))
        print(json.dumps(dict(bench="groupnorm (no SiLU)", case=name, ms=med, algorithmic_GB=3 * nbytes / 1e9,
                              GBs=3 * nbytes / med / 1e6, frac_of_hbm=3 * nbytes / med / 1e6 / hbm, peak_src=src)), flush=True)
        med, best = time_cuda(lambda: ops.bias_residual_nhwc(x, bias=b, residual=res, inplace=False))
        print(json.dumps(dict(bench="bias+residual", case=name, ms=med, algorithmic_GB=3 * nbytes / 1e9,
                              GBs=3 * nbytes / med / 1e6, frac_of_hbm=3 * nbytes / med / 1e6 / hbm, peak_src=src)), flush=True)


def bench_upconv(args):
    """Upsample2D (nearest 2x + 3x3 convolution + bias): the upsample-then-convolve form against the sub-pixel form (one 2x2
    cuDNN convolution on the low-resolution tensor + rd_phase_shuffle_bias_nhwc), on the VAE-decoder and UNet shapes of the
    fast preset.  Also times the im2col + cuBLAS GEMM form of the same 2x2 convolution (torch ops) as a probe."""
    import torch.nn.functional as F
    from rollingdepth_b200.sd2.layers import Upsample2D
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    dtype = torch.float32 if args.full else torch.float16
    if dtype == torch.float32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    for name, N, H, W, Cc in [("dec up0 12x54x96x512", 12, 54, 96, 512), ("dec up1 12x108x192x512", 12, 108, 192, 512),
                              ("dec up2 12x216x384x256", 12, 216, 384, 256), ("unet up0 48x7x12x1280", 48, 7, 12, 1280),
                              ("unet up2 48x27x48x640", 48, 27, 48, 640)]:
        torch.manual_seed(0)
        m = Upsample2D(Cc).to(dev, dtype).to(memory_format=torch.channels_last).eval()
        m.conv.fused = True
        m.fused = True
        x = torch.randn(N, Cc, H, W, device=dev, dtype=dtype).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            m.subpixel = False
            ref = m(x)
            old, _ = time_cuda(lambda: m(x))
            m.subpixel = True
            got = m(x)
            new, _ = time_cuda(lambda: m(x))
            pw = m.phase_weight()
            conv_only, _ = time_cuda(lambda: F.conv2d(x, pw, None, 1, 1))
            conv3_only, _ = time_cuda(lambda: F.conv2d(ref, m.conv.weight, None, 1, 1))
            w2 = pw.permute(0, 2, 3, 1).reshape(4 * Cc, 4 * Cc).contiguous()           # [4C out, (u, v, c)]

            def gemm_form():
                xp = F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1))                      # [N, H+2, W+2, C]
                cols = torch.cat([xp[:, u:u + H + 1, v:v + W + 1] for u in (0, 1) for v in (0, 1)], dim=-1)
                return cols.reshape(-1, 4 * Cc) @ w2.t()
            gm, _ = time_cuda(gemm_form)
            only_gemm_in = torch.randn(N * (H + 1) * (W + 1), 4 * Cc, device=dev, dtype=dtype)
            g_only, _ = time_cuda(lambda: only_gemm_in @ w2.t())
        flops36 = 2.0 * N * 4 * H * W * Cc * Cc * 9
        print(json.dumps(dict(bench="upsample2d", case=name, dtype=str(dtype), upsample_conv_bias_ms=old, subpixel_ms=new,
                              speedup=old / new, conv3x3_only_ms=conv3_only, conv3x3_TFLOPs=flops36 / conv3_only / 1e9,
                              conv2x2_only_ms=conv_only, conv2x2_TFLOPs=flops36 * 16 / 36 / conv_only / 1e9,
                              im2col_gemm_ms=gm, gemm_only_ms=g_only,
                              max_abs_diff=float((got.float() - ref.float()).abs().max()),
                              ref_absmax=float(ref.float().abs().max()))), flush=True)


def bench_energy(args):
    """Energy per call (NVML total-energy counter over a ~3 s loop) and time per call of GroupNorm, of bias+residual and of the
    two Upsample2D forms: the `fast` step runs under sw_power_cap, where a kernel that moves the same
    bytes faster does not shorten the step unless it also spends less energy."""
    import pynvml
    import torch.nn.functional as F
    from rollingdepth_b200 import ops
    from rollingdepth_b200.sd2.layers import Upsample2D
    dev = torch.device("cuda:0")
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    torch.backends.cudnn.benchmark = True

    def measure(name, fn, seconds=3.0):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        iters = max(20, int(seconds / ((time.time() - t0) / 20)))
        torch.cuda.synchronize()
        time.sleep(0.5)
        e0, t0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h), time.time()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        e1, t1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h), time.time()
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        print(json.dumps(dict(bench="energy", case=name, iters=iters, ms_per_call=(t1 - t0) / iters * 1e3,
                              mJ_per_call=(e1 - e0) / iters, avg_watts=(e1 - e0) / 1e3 / (t1 - t0), sm_mhz_after=clk)), flush=True)

    N, H, W, Cc = 12, 432, 768, 128
    x = torch.randn(N, Cc, H, W, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    res = torch.randn_like(x)
    w = torch.ones(Cc, device=dev, dtype=torch.float16)
    b = torch.zeros(Cc, device=dev, dtype=torch.float16)
    measure("groupnorm+silu dec L3 12x432x768x128", lambda: ops.group_norm_nhwc(x, 32, w, b, 1e-6, True))
    measure("bias+residual dec L3 12x432x768x128", lambda: ops.bias_residual_nhwc(x, bias=b, residual=res, inplace=False))
    del x, res
    m = Upsample2D(512).to(dev, torch.float16).to(memory_format=torch.channels_last).eval()
    m.fused = True
    m.conv.fused = True
    xu = torch.randn(12, 512, 108, 192, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for sp, tag in ((False, "nearest + 3x3"), (True, "sub-pixel 2x2 + interleave")):
            m.subpixel = sp
            measure(f"Upsample2D dec up1 12x108x192x512, {tag}", lambda: m(xu))
        a = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
        measure("cuBLAS fp16 GEMM 8192^3 (context: a tensor-bound kernel)", lambda: a @ a)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["attn", "attn_fp32", "align", "gn", "upconv", "energy"])
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--flags", type=int, nargs="+", default=[0])
    ap.add_argument("--only", default="")
    ap.add_argument("--no-sdpa", action="store_true", help="skip the torch SDPA comparators")
    a = ap.parse_args()
    {"attn": bench_attn, "attn_fp32": bench_attn_fp32, "align": bench_align, "gn": bench_gn, "upconv": bench_upconv, "energy": bench_energy}[a.what](a)
