"""Per-kernel share of a step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python tools/launch_shares.py gpurun_out/launches_r1.csv > profiles/r1_launch_shares.txt

ncu serialises launches and runs them cold-cache, so only the SHARES are meaningful (B200_PROFILING.md); the
absolute step time comes from bench.py's CUDA events.
"""
import collections
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:96]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = {h: i for i, h in enumerate(r)}
            continue
        if len(r) < len(hdr):
            continue
        if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[hdr["Metric Unit"]]
        v = float(r[hdr["Metric Value"]].replace(",", ""))
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((r[hdr["Kernel Name"]], ns))
    # one bench step ends with the renormalisation kernel: keep the LAST COMPLETE step (the first one also
    # contains cuDNN's algorithm-search launches)
    ends = [i for i, (k, _) in enumerate(rows) if "renorm_kernel" in k]
    if len(ends) >= 2:
        print(f"# step boundaries (renorm_kernel) at launches {ends}; using launches {ends[-2] + 1}..{ends[-1]}")
        rows = rows[ends[-2] + 1: ends[-1] + 1]
    tot = sum(ns for _, ns in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        a = agg[short(k)]
        a[0] += 1
        a[1] += ns
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.1f} ms of serialised kernel time")
    print(f"{'share':>7s} {'ms':>10s} {'launches':>9s} {'avg us':>9s}  kernel")
    ours = 0.0
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        mark = "*" if k.startswith("rd::") else " "
        if mark == "*":
            ours += ns
        print(f"{100 * ns / tot:6.2f}% {ns / 1e6:10.2f} {n:9d} {ns / n / 1e3:9.1f} {mark}{k}")
    print(f"# hand-written kernels of this repo (rd::*): {100 * ours / tot:.1f}% of the serialised kernel time")


if __name__ == "__main__":
    main(sys.argv[1])
