"""End-to-end GPU parity of the drop-in pipeline (`-m gpu`): reference fixtures (tiny SD2-architecture models,
CPU fp32 reference output) and the oracle pipeline on the same device."""
import os

import numpy as np
import pytest
import torch

import model_utils as mu
from oracle import pipeline as opipe

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch.device("cuda:0")


def _tiny(dev, dtype):
    from rollingdepth_b200.sd2.unet import UNet2DConditionModel
    from rollingdepth_b200.sd2.vae import AutoencoderKL
    unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1)
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
    return unet.to(dev, dtype), vae.to(dev, dtype)


def _absrel(a, b):
    d0, d1 = (a.float() + 1) / 2, (b.float() + 1) / 2
    return float(((d0 - d1).abs() / d1.clamp_min(1e-3)).mean())


def _run(pipe, frames, dil, refine, gen_dev, iters=30):
    return pipe.forward(input_frames=frames, dilations=list(dil), cap_dilation=True, snippet_lengths=[3],
                        init_infer_steps=[1], strides=[1], coalign_kwargs=dict(num_iterations=iters), refine_step=refine,
                        refine_snippet_len=3, refine_start_dilation=3, generator=gen_dev, verbose=False, max_vae_bs=4,
                        unload_snippet=False)


@pytest.mark.parametrize("tag,dil,refine", [("norefine", [1, 3], 0), ("refine", [1, 2, 3], 2)])
@pytest.mark.parametrize("snippet_batch", [1, 4])
def test_pipeline_fp32_matches_reference_fixture(golden_dir, tag, dil, refine, snippet_batch):
    """Same weights (name-keyed init), same frames, same text embedding, same CPU noise generator."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    dev = _dev()
    # the fixture is a CPU fp32 run: keep cuDNN / cuBLAS in true fp32 (torch allows TF32 convolutions by default)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(golden_dir, "models_pipeline.npz"))
    unet, vae = _tiny(dev, torch.float32)
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=snippet_batch)
    pipe.empty_text_embed = torch.from_numpy(g["pipe_text"]).to(dev)
    frames = torch.from_numpy(g["pipe_frames"])

    res = _run(pipe, frames, dil, refine, torch.Generator().manual_seed(0))   # the CPU generator of the reference run
    for i, s in enumerate(res.snippet_ls):
        ref = torch.from_numpy(g[f"pipe_{tag}_snip{i}"])
        assert s.shape == ref.shape
        assert float((s - ref).abs().max()) < 2e-4
    co = torch.from_numpy(g[f"pipe_{tag}_coaligned"])
    dp = torch.from_numpy(g[f"pipe_{tag}_depth_pred"])
    assert res.depth_pred.shape == dp.shape and res.depth_coaligned.shape == co.shape
    assert float((res.input_rgb - torch.from_numpy(g[f"pipe_{tag}_rgb"])).abs().max()) < 1e-6
    assert _absrel(res.depth_coaligned, co) <= 1e-3
    assert _absrel(res.depth_pred, dp) <= 1e-3      # north star: final depth AbsRel difference <= 1e-3


def test_pipeline_fp16_matches_oracle_on_device():
    """fp16: product pipeline (tcgen05 attention, fused DDIM/aligner kernels, batched snippets) against the
    oracle's op-for-op restatement run with library attention on the same GPU."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    from rollingdepth_b200.sd2.layers import TorchSDPAProcessor
    dev = _dev()
    unet, vae = _tiny(dev, torch.float16)
    g = torch.Generator().manual_seed(3)
    frames = torch.rand(1, 12, 3, 48, 64, generator=g) * 2 - 1
    frames = torch.nn.functional.avg_pool2d(frames[0], 5, 1, 2)[None]
    text = torch.randn(1, 2, 32, generator=g).to(dev).half()
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=3)
    pipe.empty_text_embed = text
    res = _run(pipe, frames, [1, 3], 2, torch.Generator(device=dev).manual_seed(0), iters=20)
    unet.set_attn_processor(TorchSDPAProcessor())
    ora = opipe.forward(unet, vae, mu.DDIM_CFG, text, frames, [1, 3], cap_dilation=True, coalign_kwargs=dict(num_iterations=20),
                        refine_step=2, refine_start_dilation=3, generator=torch.Generator(device=dev).manual_seed(0))
    for a, b in zip(res.snippet_ls, ora.snippet_ls):
        assert float((a.float() - b.float().cpu()).abs().max()) < 2e-2
    assert _absrel(res.depth_pred, ora.depth_pred.cpu()) <= 3e-3   # fp16 noise floor of the tiny random model
    assert torch.isfinite(res.depth_pred.float()).all()


def test_output_modes_agree_and_results_are_fresh_by_default(golden_dir):
    """Default outputs are fresh pageable CPU tensors (the reference's `.cpu()` semantics) that survive the next call;
    `output_mode = "pinned_views"` returns the same values as views of the pipeline's staging buffers."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline
    from rollingdepth_b200.scheduler import DDIMSchedule
    dev = _dev()
    g = np.load(os.path.join(golden_dir, "models_pipeline.npz"))
    unet, vae = _tiny(dev, torch.float32)
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=4)
    pipe.empty_text_embed = torch.from_numpy(g["pipe_text"]).to(dev)
    frames = torch.from_numpy(g["pipe_frames"])
    a = _run(pipe, frames, [1, 3], 0, torch.Generator().manual_seed(0))
    keep = a.depth_pred.clone()
    assert not a.depth_pred.is_pinned() and not a.snippet_ls[0].is_pinned()
    b = _run(pipe, frames.flip(1), [1, 3], 0, torch.Generator().manual_seed(0))     # different input, same buffers
    assert torch.equal(a.depth_pred, keep), "results of an earlier call were overwritten"
    assert not torch.equal(a.depth_pred, b.depth_pred)
    pipe.output_mode = "pinned_views"
    c = _run(pipe, frames, [1, 3], 0, torch.Generator().manual_seed(0))
    assert c.depth_pred.is_pinned()
    # two separate forward passes: equal up to the order of the shared-memory double atomics in the GroupNorm statistics
    close = lambda x, y: float((x - y).abs().max()) <= 1e-4
    assert close(c.depth_pred, a.depth_pred) and close(c.depth_coaligned, a.depth_coaligned)
    for x, y in zip(c.snippet_ls, a.snippet_ls):
        assert close(x, y)
