"""CPU checks of bench.py's host logic: the exact unit counts it extrapolates with, the
JSON contract of the `--impl reference` arm (run here on a tiny stand-in model so that it finishes in seconds), and the
rule that under torchrun only rank 0 of that arm works and prints."""
import importlib.util
import io
import json
import os
import subprocess
import sys
import types
from contextlib import redirect_stdout

import pytest
import torch

import model_utils as mu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("rd_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_call_counts_match_survey(bench):
    """fast = 148 + 100 = 248 snippets (744 decodes, 150 encodes); full = 378 snippets + 1432
    refinement windows, 1134 snippet-frame decodes + 150 final decodes, 2 x 150 encodes; cfg1 = 14 + 8 snippets."""
    dil, c = bench.call_counts("fast")
    assert dil == [1, 25] and c == dict(unet=248, decode=744, encode=150, align_it=2000)
    dil, c = bench.call_counts("full")
    assert dil == [1, 10, 25]
    assert c == dict(unet=378 + 1432, decode=3 * 378 + 150, encode=300, align_it=2000)
    dil, c = bench.call_counts("cfg1")
    assert dil == [1, 4] and c["unet"] == 14 + 8 and c["decode"] == 66 and c["encode"] == 16
    dil, c = bench.call_counts("paper")          # cap_dilation off: the fixed dilations of the paper
    assert dil == [1, 10, 25] and c["unet"] == 378 + 1432


def test_call_counts_agree_with_product_scheduler(bench):
    """The counts are derived from the oracle's scheduler; the product's own scheduler must give the same numbers."""
    from rollingdepth_b200.pipeline import RollingDepthPipeline as P
    for preset in ("fast", "full", "cfg1"):
        n, _H, _W, dil, refine, _dt, cap = bench.PRESETS[preset]
        d = [P.cap_max_dilation(n, 3, x, False) for x in dil] if cap else list(dil)
        n_unet = sum(len(P.get_snippet_indice(0, [0], n, 3, x, x, 1)) for x in d)
        if refine:
            start = P.cap_max_dilation(n, 3, 6, False) if cap else 6
            n_unet += sum(len(P.get_snippet_indice(i, [0] * refine, n, 3, start, 1, 1)) for i in range(refine))
        assert n_unet == bench.call_counts(preset)[1]["unet"], preset


def test_workload_config_names_the_preset(bench):
    cfg = bench.workload_config("fast", 1)
    assert "fast" in cfg["workload"] and "768x432" in cfg["workload"] and "float16" in cfg["workload"]
    assert cfg["frames"] == 150 and cfg["dilations"] == [1, 25] and cfg["parallelism"] == "single GPU"
    assert "model" not in cfg
    assert bench.workload_config("full", 8)["parallelism"] == "snippet-sharded x8"


def _tiny_models(device, dtype, seed=0):
    from rollingdepth_b200.sd2.unet import UNet2DConditionModel
    from rollingdepth_b200.sd2.vae import AutoencoderKL
    cfg = dict(mu.TINY_UNET, cross_attention_dim=1024)          # bench.py feeds a [1, 2, 1024] text embedding
    unet = mu.deterministic_init_(UNet2DConditionModel(**cfg).eval(), seed=1)
    vae = mu.deterministic_init_(AutoencoderKL(**{k: v for k, v in mu.TINY_VAE.items() if not k.endswith("block_types")}).eval(), seed=2)
    return unet.to(device, dtype), vae.to(device, dtype)


def test_reference_arm_json_contract(bench, monkeypatch):
    """`bench.py --impl reference`: one JSON line with the arm's keys; value = frames / extrapolated seconds of the
    sampled units; e2e repeats the line's own value with zero copy bytes (the CPU path has no device)."""
    monkeypatch.setattr(bench, "build_models", _tiny_models)
    monkeypatch.setitem(bench.PRESETS, "fast", (12, 32, 48, [1, 3], 0, torch.float16, True))
    monkeypatch.delenv("RANK", raising=False)
    args = types.SimpleNamespace(preset="fast", steps=1, warmup=0, gpus=1, cpu_full=False, snippet_batch=16, max_vae_bs=36,
                                 channels_last=1, fused_norm=1, cudnn_benchmark=1, subpixel=1)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] > 0
    assert set(cb["per_unit_seconds"]) == {"encode", "unet", "decode", "align_it"}
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["ms_per_step"] - 1e3 * 12 / d["value"]) < 1e-6 * d["ms_per_step"]
    # the extrapolation is exactly counts x per-unit seconds
    _dil, counts = bench.call_counts("fast")
    total = sum(counts[k] * cb["per_unit_seconds"][k] for k in counts)
    assert abs(cb["seconds_per_video"] - total) < 1e-9 * total


def test_reference_arm_other_ranks_exit_without_work():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 and print nothing."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_cuda():
    """No CPU fallback: the product arm must not produce a bench line on a box without a GPU."""
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert '"metric"' not in r.stdout
