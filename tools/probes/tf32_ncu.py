import os, sys, torch
sys.path.insert(0, "/root/repo")
from rollingdepth_b200 import ops, _lib as L
dev = torch.device("cuda:0")
S, H = 15552, 5
q, k, v = [torch.randn(1, S, H * 64, device=dev).view(1, S, H, 64) for _ in range(3)]
out = torch.empty_like(q)
for _ in range(4):
    ops.xattn_fwd(q, k, v, out=out)
torch.cuda.synchronize()
print("ok")
