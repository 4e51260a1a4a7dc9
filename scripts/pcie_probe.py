"""PCIe copy bandwidth probe (GPU box): contiguous vs 2-D strided pinned copies, both directions, and concurrently."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from seqikpy_b200 import _native as N
lib = N.load_library()
n_chain, n_frame = 6000, 1000
h = torch.empty((n_chain, n_frame, 27), dtype=torch.float32, pin_memory=True)
d = torch.empty((n_chain, n_frame, 27), dtype=torch.float32, device="cuda")
h2 = torch.empty((n_chain, n_frame, 15), dtype=torch.float32, pin_memory=True)
d2 = torch.empty((n_chain, n_frame, 15), dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def c2d(dst, src, row, t0, t1, direction, stream):
    off = 4 * row * t0
    N.check(lib.seqik_memcpy2d_async(dst.data_ptr() + off, 4 * row * n_frame, src.data_ptr() + off, 4 * row * n_frame, 4 * row * (t1 - t0), n_chain, direction, stream.cuda_stream), "c")

cur = torch.cuda.current_stream()
gb = h.numel() * 4 / 1e9
print("D2H contiguous 648 MB: %.1f GB/s" % (gb / t(lambda: h.copy_(d, non_blocking=True)) * 1e3))
print("H2D contiguous 648 MB: %.1f GB/s" % (gb / t(lambda: d.copy_(h, non_blocking=True)) * 1e3))
for chunks in (1, 4, 8, 16):
    def f():
        for k in range(chunks):
            c2d(h, d, 27, n_frame * k // chunks, n_frame * (k + 1) // chunks, 2, cur)
    print("D2H 2-D strided in %2d frame chunks: %.1f GB/s" % (chunks, gb / t(f) * 1e3))
def both():
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
ms = t(both)
print("concurrent D2H 648 MB + H2D 360 MB: %.2f ms -> D2H %.1f GB/s, H2D %.1f GB/s" % (ms, gb / ms * 1e3, h2.numel() * 4 / 1e9 / ms * 1e3))
