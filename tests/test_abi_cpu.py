"""CPU checks of the C-ABI boundary and of the host-side sharding logic (no GPU compute)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(rd_[a-z0-9_]+)\s*\(", src)) - {"rd_status", "rd_dtype"}


def test_library_exports_every_declared_symbol():
    from rollingdepth_b200 import _lib as L
    from rollingdepth_b200 import build
    build.build()
    declared = _header_functions()
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rd_version() == 1


def test_sass_contains_blackwell_instructions():
    """The attention kernel must be a tcgen05/TMA kernel, not a recompiled mma.sync one."""
    from rollingdepth_b200 import build
    lib = build.build()
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass


def test_compute_entry_points_fail_loudly_without_cuda():
    from rollingdepth_b200 import ops, _lib as L
    from rollingdepth_b200.depth_aligner import DepthAligner
    q = torch.zeros(1, 8, 1, 64, dtype=torch.float16)
    with pytest.raises(L.RdError):
        ops.xattn_fwd(q, q, q)
    with pytest.raises(L.RdError):
        DepthAligner(device=torch.device("cpu")).run([torch.zeros(2, 3, 1, 24, 24)], [1])
    if not torch.cuda.is_available():
        assert L.lib().rd_device_arch() < 0
        assert b"CUDA" in L.lib().rd_last_error()


def test_problem_init_and_argument_checks():
    import ctypes as C
    from rollingdepth_b200 import _lib as L
    pb = L.CoalignProblem()
    pb.G, pb.w, pb.N, pb.H, pb.W, pb.border, pb.factor = 2, 3, 150, 432, 768, 2, 10
    pb.d[0], pb.d[1] = 1, 25
    assert L.lib().rd_coalign_problem_init(C.byref(pb)) == 0
    assert (pb.n[0], pb.n[1], pb.n_total) == (148, 100, 248)
    assert (pb.Hs, pb.Ws, pb.P, pb.P_pad) == (43, 77, 3311, 3312)
    pb.border = 0      # the reference's [border:-border] slice is empty for 0
    assert L.lib().rd_coalign_problem_init(C.byref(pb)) == -1
    assert b"border" in L.lib().rd_last_error()
    pb.border, pb.d[1] = 2, 80
    assert L.lib().rd_coalign_problem_init(C.byref(pb)) == -1   # dilation too large for the clip


def test_block_partition_is_contiguous_with_minimal_makespan():
    """Slots of ceil(n / world) units: contiguous, complete, the largest block as small as a balanced partition's,
    and every rank's block starts at rank * cap -- which is what lets one in-place all-gather assemble the result."""
    from rollingdepth_b200.sharding import Shard
    for n in (0, 1, 7, 9, 100, 148, 248, 378):
        for world in (1, 2, 3, 4, 8):
            blocks = [Shard.block_of(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            cap = Shard.cap_of(n, world)
            assert max(sizes) == cap == -(-n // world)
            assert all(a == min(n, r * cap) for r, (a, _) in enumerate(blocks))
    assert Shard.block_of(248, 3, 8) == (93, 124)      # the `fast` preset's 148 + 100 snippets: 31 per rank, no remainder


def test_balanced_chunks():
    from rollingdepth_b200.sharding import balanced_chunks
    assert balanced_chunks(19, 16) == [(0, 10), (10, 19)]           # not 16 + 3
    assert balanced_chunks(57, 36) == [(0, 29), (29, 57)]
    assert balanced_chunks(16, 16) == [(0, 16)] and balanced_chunks(0, 16) == []
    for n in range(1, 80):
        for m in (1, 3, 16, 36):
            ch = balanced_chunks(n, m)
            assert ch[0][0] == 0 and ch[-1][1] == n and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
            sizes = [b - a for a, b in ch]
            assert max(sizes) <= m and max(sizes) - min(sizes) <= 1 and len(ch) == -(-n // m)


def test_aligner_geometry_matches_reference(golden_dir):
    """Index / shape bookkeeping of `DepthAligner.run` (depth_aligner.py:68-100): the product's `_problem`
    (rd_coalign_problem_init: host C code of the library, no device) and `create_triplet_indices` against what the REFERENCE
    hands to its `optimize` -- sequence length, cropped + subsampled sizes, snippet counts, triplet index tables -- for 116
    combinations of frame count, window, dilations, resolution, border and factor (tests/golden/aligner_geometry.json)."""
    import json
    import zlib
    import torch
    from rollingdepth_b200.depth_aligner import DepthAligner
    cases = json.load(open(os.path.join(golden_dir, "aligner_geometry.json")))
    assert len(cases) >= 116
    for c in cases:
        al = DepthAligner(device=torch.device("cuda", 0), factor=c["factor"], border=c["border"])
        snips = [torch.empty((n, c["w"], 1, c["H"], c["W"]), device="meta") for n in c["n"]]
        pb = al._problem(snips, c["dilations"])
        assert pb.N == c["N"] and pb.G == len(c["dilations"]) and pb.w == c["w"], c
        row0 = 0
        for g, small in enumerate(c["small"]):
            assert [pb.n[g], pb.w, 1, pb.Hs, pb.Ws] == small, (c, g)
            assert pb.row0[g] == row0
            row0 += pb.n[g]
            idx = al.create_triplet_indices(pb.N, pb.d[g] - 1, pb.w).tolist()
            flat = [x for win in idx for x in win]
            got = dict(n=len(idx), first=idx[0], last=idx[-1], index_sum=sum(flat),
                       crc=zlib.crc32(",".join(map(str, flat)).encode()) & 0xFFFFFFFF)
            assert got == c["idx"][g], (c, g)
        assert pb.n_total == row0 and pb.P == pb.Hs * pb.Ws and pb.P_pad % 8 == 0 and 0 <= pb.P_pad - pb.P < 8


def test_product_snippet_scheduler_matches_reference(golden_dir):
    """The PRODUCT copy of get_snippet_indice / cap_max_dilation (rollingdepth_pipeline.py:465-515) against all
    reference-generated cases (integer work: exact)."""
    import json
    from rollingdepth_b200.pipeline import RollingDepthPipeline as P
    cases = json.load(open(os.path.join(golden_dir, "snippets.json")))
    assert len(cases) >= 96
    for c in cases:
        if c["kind"] == "init":
            capped = P.cap_max_dilation(seq_len=c["seq_len"], snippet_len=c["snippet_len"], dilation=c["dilation"], verbose=False)
            assert capped == c["capped"], c
            idx = P.get_snippet_indice(0, [999], c["seq_len"], c["snippet_len"], capped, capped, 1)
        else:
            idx = P.get_snippet_indice(c["i_step"], list(range(c["total"])), c["seq_len"], 3, c["start"], 1, 1)
        assert idx == c["idx"], c


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["RD_ROOT"])
from rollingdepth_b200.sharding import Shard
from oracle import schedule as osch
dist.init_process_group("gloo")
sh = Shard.from_dist()
# (1) in-place all-gather of per-rank slots reproduces the full tensor in global order (uneven and empty blocks)
for total in (7, 1, 8):
    full = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
    lo, hi = sh.block(total)
    buf, mine = sh.gather_buffer(total, (3,), torch.float32, "cpu")
    assert buf.shape[0] == sh.world * sh.cap(total) and mine.shape[0] == hi - lo
    mine.copy_(full[lo:hi])
    got = sh.gather_(buf, total)
    assert torch.equal(got, full), (got, full)
# (2) snippet-sharded "denoise" over the flattened (dilation, snippet) list: identical to 1 rank
N, n = 30, 3
flat = []
for d in (1, 4):
    flat += osch.get_snippet_indice(0, 1, N, n, d, d, 1)
frames = torch.arange(N, dtype=torch.float32)
def work(ix): return torch.stack([frames[i].sum() * 2 + 1 for i in map(torch.tensor, ix)]) if ix else torch.zeros(0)
lo, hi = sh.block(len(flat))
buf, mine = sh.gather_buffer(len(flat), (), torch.float32, "cpu")
mine.copy_(work(flat[lo:hi]))
assert torch.equal(sh.gather_(buf, len(flat)), work(flat))
# (3) sum / min / max all-reduce and broadcast
acc = torch.zeros(N)
for k in range(lo, hi):
    for j, f in enumerate(flat[k]):
        acc[f] += float(k + j)
sh.all_reduce_(acc)
ref_a = torch.zeros(N)
for k in range(len(flat)):
    for j, f in enumerate(flat[k]):
        ref_a[f] += float(k + j)
assert torch.equal(acc, ref_a)
mn = torch.tensor([float(sh.rank + 3)]); sh.all_reduce_(mn, "min"); assert float(mn) == 3.0
mx = torch.tensor([float(sh.rank + 3)]); sh.all_reduce_(mx, "max"); assert float(mx) == 2.0 + sh.world
b = torch.full((2,), float(sh.rank)); sh.broadcast_(b); assert float(b.sum()) == 0.0
dist.destroy_process_group()
print("rank", sh.rank, "ok")
'''


def test_sharding_collectives_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, RD_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_launch_share_tool_on_a_synthetic_ncu_csv(tmp_path, capsys):
    """tools/launch_shares.py: shares of the LAST complete step (steps end with the renormalisation kernel)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("launch_shares", os.path.join(ROOT, "tools", "launch_shares.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"'
    def row(i, name, ns):
        return f'"{i}","1","python","h","{name}","1","7","(256, 1, 1)","(1, 1, 1)","0","10.0","s","gpu__time_duration.sum","ns","{ns}"'
    names = [("void rd::xattn_fwd_kernel<__half, 0>(int)", 1000), ("cudnn_conv(int)", 3000), ("void rd::renorm_kernel<__half>(int)", 10)]
    lines = ["==PROF== header line", hdr]
    i = 0
    for rep in range(2):                        # two complete steps; the second has a slower attention launch
        for n, ns in names:
            lines.append(row(i, n, ns * (2 if (rep == 1 and "xattn" in n) else 1)))
            i += 1
    p = tmp_path / "launches.csv"
    p.write_text("\n".join(lines) + "\n")
    mod.main(str(p))
    out = capsys.readouterr().out
    assert "using launches 3..5" in out
    assert "39.92%" in out and "*rd::xattn_fwd_kernel" in out       # 2000 / 5010
