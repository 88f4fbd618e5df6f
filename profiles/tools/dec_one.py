import sys, torch
sys.path.insert(0, "/root/repo")
from pqdet_b200 import _ops
C, size, nB = 10, 608, 256
raws = [torch.randn((nB, 3 * (5 + C), size // s, size // s), device="cuda") for s in (16, 8)]
for _ in range(3):
    _ops.decode_levels(raws, C, (16, 8))
torch.cuda.synchronize()
