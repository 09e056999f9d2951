"""Multi-GPU (snippet-sharded) result invariance: 2 / 4 / 8 ranks over NCCL vs 1 rank, same inputs (`-m gpu`; each case
is skipped when the box has fewer GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RD_ROOT"]); sys.path.insert(0, os.path.join(os.environ["RD_ROOT"], "tests"))
import model_utils as mu
from rollingdepth_b200.pipeline import RollingDepthPipeline
from rollingdepth_b200.scheduler import DDIMSchedule
from rollingdepth_b200.sd2.unet import UNet2DConditionModel
from rollingdepth_b200.sd2.vae import AutoencoderKL
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = False
unet = mu.deterministic_init_(UNet2DConditionModel(**mu.TINY_UNET).eval(), seed=1).to(dev).half()
vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2).to(dev).half()
g = torch.Generator().manual_seed(3)
N = 4 * world + 6
frames = torch.rand(1, N, 3, 48, 64, generator=g) * 2 - 1
frames = torch.nn.functional.avg_pool2d(frames[0], 5, 1, 2)[None]
text = torch.randn(1, 2, 32, generator=g).to(dev).half()
def run(shard, **kw):
    # batch sizes of 1 keep the cuDNN / cuBLAS problem shapes identical between the sharded and the single run
    pipe = RollingDepthPipeline(unet, vae, DDIMSchedule(**mu.DDIM_CFG), snippet_batch=1)
    pipe.empty_text_embed = text
    if shard: pipe.enable_sharding(**kw)
    out = pipe.forward(input_frames=frames, dilations=[1, 3], cap_dilation=True, snippet_lengths=[3], init_infer_steps=[1],
                       strides=[1], coalign_kwargs=dict(num_iterations=40), refine_step=2, refine_snippet_len=3,
                       refine_start_dilation=3, generator=torch.Generator().manual_seed(0), verbose=False, max_vae_bs=1,
                       unload_snippet=False)
    return out, pipe.last_alignment
single, al1 = run(False)
for mode in ("replicated", "allreduce"):
    sharded, al2 = run(True, output_ranks=None, coalign_mode=mode)      # every rank returns the complete result
    assert len(sharded.snippet_ls) == len(single.snippet_ls)
    for a, b in zip(single.snippet_ls, sharded.snippet_ls):             # the complete snippet list, as the reference returns
        assert a.shape == b.shape and torch.equal(a, b), float((a.float() - b.float()).abs().max())
    for a, b in zip(al1[0] + al1[1], al2[0] + al2[1]):                  # scales and translations: bit-identical
        assert torch.equal(a, b), (mode, float((a - b).abs().max()))
    for h1, h2 in zip(al1[2], al2[2]):
        assert abs(h1[0] - h2[0]) <= 1e-6 * abs(h1[0]) and h1[1] == h2[1] and h1[2] == h2[2], (mode, h1, h2)
    assert torch.equal(single.depth_coaligned, sharded.depth_coaligned), mode
    assert torch.equal(single.depth_pred, sharded.depth_pred), mode
    assert torch.equal(single.input_rgb, sharded.input_rgb)
    print(f"rank {rank}: {mode}: bit-identical to the single-GPU run", flush=True)
# default: only rank 0 returns (and copies) results
part, _ = run(True)
if rank == 0:
    assert torch.equal(part.depth_pred, single.depth_pred) and len(part.snippet_ls) == 2
else:
    assert part.depth_pred is None and part.snippet_ls is None and part.input_rgb is None
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok", flush=True)
'''


@pytest.mark.parametrize("world", [2, 4, 8])
def test_snippet_sharding_matches_single_gpu_bit_for_bit(tmp_path, world):
    """Sharded over `world` GPUs vs one GPU on the same inputs: decoded snippets (complete list on every rank), scales,
    translations, merged and refined depth are bit-identical, with the replicated optimiser and with the frame-sharded
    all-reduce optimiser (rd_coalign_begin / pass / exchange / end)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, RD_ROOT=ROOT)
    # one retry on another port: a rendezvous / NCCL bring-up hiccup on a fresh box is not a parity failure
    # (the numerical assertions live in the worker and would fail identically both times)
    for port in ("29541", "29547"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                           capture_output=True, text=True, env=env, timeout=900)
        if r.returncode == 0 or "AssertionError" in r.stdout + r.stderr:
            break
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == world
