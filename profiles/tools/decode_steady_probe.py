import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops
def ev(fn, reps=7, inner=6):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner): fn()
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / inner)
    return sorted(ts)[len(ts) // 2]
B, C = 256, 20
raws = [torch.randn((B, 75, 512 // s, 512 // s), device="cuda") for s in (32, 16, 8)]
t = ev(lambda: _ops.decode_levels(raws, C, (32, 16, 8))) * 1e3
nbytes = 2 * sum(r.numel() for r in raws) * 4
print("lib %s: decode_levels %.1f us = %.0f GB/s of 2R" % (os.environ.get("PQDET_B200_LIB", "default"), t, nbytes / t / 1e3))
x = torch.randn((B, 80, 64, 64), device="cuda"); w = torch.randn((75, 80), device="cuda") * 0.03; bias = torch.randn((75,), device="cuda") * 0.1
o = torch.empty((B, 64, 64, 3, 25), device="cuda")
t = ev(lambda: _ops.head_conv_decode(x, w, bias, C, 8.0, out=o, rows_total=64 * 64 * 3, row_offset=0)) * 1e3
print("   head conv stride-8 level: %.1f us = %.0f GB/s" % (t, (x.numel() * 4 + o.numel() * 4) / t / 1e3))
