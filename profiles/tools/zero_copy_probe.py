"""Probe: fused decode+NMS reading the raw heads straight from pinned (UVA-mapped) host memory vs staging them
with an H2D copy first.  python profiles/tools/zero_copy_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pqdet_b200 import _ops, _lib, synth

B, C, SIZE = 1024, 20, 512
STR = (32, 16, 8)
dev = torch.device("cuda", 0)
heads = synth.make_heads(B, C, SIZE, "sparse", seed=0, device=dev)
host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in heads]
for h, t in zip(host, heads):
    h.copy_(t)
torch.cuda.synchronize()
orig = torch.tensor([512.0, 512.0], device=dev)
hd, keep = _ops.make_heads(heads, STR, C, (SIZE, SIZE), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda")
out = _ops.alloc_fused_outputs(B, 2048, False, dev)
_ops.decode_nms_fused(hd, keep, 2048, False, out=out)
torch.cuda.synchronize()
ref_det, ref_meta = out[0].clone(), out[2].clone()

# the same descriptor with host pointers
hz, keepz = _ops.make_heads(heads, STR, C, (SIZE, SIZE), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda")
for i, h in enumerate(host):
    hz.raw[i] = h.data_ptr()
out2 = _ops.alloc_fused_outputs(B, 2048, False, dev)

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n

def res():
    _ops.decode_nms_fused(hd, keep, 2048, False, out=out)
t = timeit(res, 20)
print("device-resident kernel: %.3f ms -> %.0f img/s" % (t * 1e3, B / t))
def zc():
    _ops.decode_nms_fused(hz, keepz, 2048, False, out=out2)
t = timeit(zc)
print("zero-copy kernel, device outputs: %.3f ms -> %.0f img/s" % (t * 1e3, B / t))
hh, hkeep = _ops.make_heads_host(host, STR, C, (SIZE, SIZE), torch.tensor([512.0, 512.0]), "voc", 0.1, 0.45, "auto_cuda", "tv_cuda")
hout = _ops.alloc_host_outputs(B, 2048, False, dev)
def zch():
    _ops.decode_nms_host(hh, hkeep, 2048, False, dev, out=hout)
t = timeit(zch)
print("zero-copy kernel, host outputs: %.3f ms -> %.0f img/s" % (t * 1e3, B / t))
print("host outputs identical:", torch.equal(hout[2][:3 * B], ref_meta[:3 * B].cpu()) and
      all(torch.equal(hout[0][b, :int(ref_meta[b])], ref_det[b, :int(ref_meta[b])].cpu()) for b in range(0, B, 37)))
cnt = ref_meta[:B]
same = torch.equal(out2[2][:3 * B], ref_meta[:3 * B])
ok = same and all(torch.equal(out2[0][b, :int(cnt[b])], ref_det[b, :int(cnt[b])]) for b in range(0, B, 37))
print("identical to device-resident run:", ok)

dev_in = [torch.empty_like(t) for t in heads]
def h2d():
    for d, s in zip(dev_in, host):
        d.copy_(s, non_blocking=True)
t = timeit(h2d)
nb = sum(x.numel() * 4 for x in heads)
print("full H2D: %.3f ms  %.1f GB/s" % (t * 1e3, nb / t / 1e9))

# objectness planes only via strided copy (cudaMemcpy2D under the hood)
stag = [torch.empty((B, 3, t.shape[2], t.shape[3]), device=dev) for t in heads]
def h2d_obj():
    for s, h in zip(stag, host):
        v = h.view(B, 3, 5 + C, h.shape[2], h.shape[3])[:, :, 4]
        s.copy_(v, non_blocking=True)
t = timeit(h2d_obj)
nb2 = sum(x.numel() * 4 for x in stag)
print("objectness planes H2D (strided): %.3f ms  %.1f GB/s" % (t * 1e3, nb2 / t / 1e9))
