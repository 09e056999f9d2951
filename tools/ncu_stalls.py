"""Summarise an `ncu --page source --csv` dump: stall mix and the hottest SASS lines of the first kernel."""
import csv
import sys


def main(path, topn=30):
    rows = list(csv.reader(open(path)))
    # sections: "Kernel Name" line, header line, data lines
    i = 0
    while i < len(rows) and not (rows[i] and rows[i][0] == "Address"):
        i += 1
    hdr = rows[i]
    idx = {h: k for k, h in enumerate(hdr)}
    data = []
    for r in rows[i + 1:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            data.append(r)
    f = lambda r, c: float(r[idx[c]] or 0)
    tot = sum(f(r, "# Samples") for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {c: sum(f(r, c) for r in data) for c in stalls}
    print(f"samples {tot:.0f}, sass lines {len(data)}")
    for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
        print(f"  {c:24s} {100 * v / max(tot, 1):5.1f}%")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
        st = sorted(((c, f(r, c)) for c in stalls), key=lambda kv: -kv[1])[:2]
        print(f"{f(r, '# Samples'):7.0f} {f(r, 'Instructions Executed'):10.0f}  {r[idx['Source']][:64]:64s} "
              + " ".join(f"{c[6:]}={v:.0f}" for c, v in st if v))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
