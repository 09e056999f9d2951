"""Kernel-time breakdown of one pipeline step with torch.profiler (diagnostic, not a benchmark)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "fast"
    frames_n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    dev = torch.device("cuda:0")
    n, H, W, dil, refine, dtype, cap = bench.PRESETS[preset]
    torch.backends.cudnn.benchmark = bool(int(os.environ.get("RD_CUDNN_BENCH", "1")))
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    unet, vae = bench.build_models(dev, dtype)
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(), snippet_batch=int(os.environ.get("RD_SB", "8")),
                                channels_last=bool(int(os.environ.get("RD_CL", "1"))),
                                fused_norm=bool(int(os.environ.get("RD_FN", "1"))))
    pipe.empty_text_embed = torch.randn(1, 2, 1024).to(dev, dtype)
    frames = bench.synth_video(frames_n, H, W).to(dev)

    def step():
        return pipe.forward(input_frames=frames, dilations=list(dil), cap_dilation=cap, snippet_lengths=[3],
                            init_infer_steps=[1], strides=[1], coalign_kwargs=None, refine_step=refine, refine_snippet_len=3,
                            refine_start_dilation=6, generator=torch.Generator(device=dev).manual_seed(0), verbose=False,
                            max_vae_bs=int(os.environ.get("RD_VAE_BS", "12")), unload_snippet=False)
    step()
    torch.cuda.synchronize()
    import time
    t0 = time.time(); step(); torch.cuda.synchronize(); print("step wall s:", time.time() - t0)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))


if __name__ == "__main__":
    main()
