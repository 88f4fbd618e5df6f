import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops
C, B, cin, s, size = int(sys.argv[1]), int(sys.argv[2]), 352, 32, 608
ch = 3 * (5 + C)
f = torch.randn((B, cin, size // s, size // s), device="cuda")
w = torch.randn((ch, cin), device="cuda") / cin ** 0.5
b = torch.randn((ch,), device="cuda") * 0.1
for _ in range(3):
    _ops.head_conv_decode(f, w, b, C, float(s))
torch.cuda.synchronize()
