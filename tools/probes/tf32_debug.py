"""Diagnostics for the 3xTF32 attention kernel (special inputs that isolate the operand paths)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rollingdepth_b200 import ops, _lib as L
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(q, k, v, tag):
    out = torch.full_like(q, float("nan"))
    ops.xattn_fwd(q, k, v, out=out, flags=L.XATTN_F32(2))
    ref = torch.empty_like(q)
    ops.xattn_fwd(q, k, v, out=ref, flags=L.XATTN_F32(1))
    torch.cuda.synchronize()
    d = (out - ref).abs()
    print(f"{tag}: max|diff| {float(d.max()):.3e}  nan {int(torch.isnan(out).sum())}  out[0,0,0,:6] {out[0,0,0,:6].tolist()}  ref {ref[0,0,0,:6].tolist()}")
    print(f"      out[0,1,0,30:36] {out[0,1,0,30:36].tolist()} ref {ref[0,1,0,30:36].tolist()}")
for S in (64, 128, 200):
    q = torch.randn(1, S, 1, 64, device=dev); k = torch.randn(1, S, 1, 64, device=dev); v = torch.randn(1, S, 1, 64, device=dev)
    run(q, k, torch.ones_like(v), f"S={S} V=1 (expect 1)")
    run(torch.zeros_like(q), k, v, f"S={S} Q=0 (expect column means of V)")
    chan = torch.arange(64, device=dev, dtype=torch.float32).view(1, 1, 1, 64).expand(1, S, 1, 64).contiguous()
    run(torch.zeros_like(q), k, chan, f"S={S} Q=0, V[:, c] = c (expect c)")
    keyid = torch.arange(S, device=dev, dtype=torch.float32).view(1, S, 1, 1).expand(1, S, 1, 64).contiguous()
    run(torch.zeros_like(q), k, keyid, f"S={S} Q=0, V[j, :] = j (expect (S-1)/2)")
    run(q, k, v, f"S={S} random")
