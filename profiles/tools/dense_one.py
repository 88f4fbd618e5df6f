import sys, torch
sys.path.insert(0, "/root/repo")
from pqdet_b200 import _ops, synth
B, C, size = 64, 10, 608
dev = torch.device("cuda")
heads = synth.make_heads(B, C, size, "dense", seed=0, device=dev)
orig = torch.tensor([480.0, 480.0], device=dev)
h, keep = _ops.make_heads(heads, (32, 16, 8), C, (size, size), orig, "visdrone", 0.1, 0.45, "auto_cuda", "tv_cuda")
for _ in range(3):
    _ops.nms_general(heads_t=h, keep_alive=keep, n_images=B, max_det=8192, cand_capacity=1 << 21)
torch.cuda.synchronize()
