"""Host-side ceilings of the int8 transport conversions on this box (no GPU work): GB/s of int32 touched."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import xsqueezeit_b200 as xb
L = xb.lib()
n = 1 << 30  # 4 GiB of int32
src = torch.empty(n, dtype=torch.int32, pin_memory=True); src.fill_(3)
dst = torch.empty(n, dtype=torch.int8, pin_memory=True); dst.zero_()
print("threads", L.xsi_host_threads())
for chunk in (n, 1 << 24, 1 << 22):
    best = 0
    for rep in range(3):
        t = time.perf_counter()
        for a in range(0, n, chunk):
            L.xsi_host_narrow_i32_i8(src.data_ptr() + 4 * a, dst.data_ptr() + a, chunk)
        best = max(best, 4 * n / (time.perf_counter() - t) / 1e9)
    print("narrow, calls of %d Mi elements: %.1f GB/s" % (chunk >> 20, best))
ln = np.array([n], np.uint32)
for chunk in (n, 1 << 24):
    best = 0
    for rep in range(3):
        t = time.perf_counter()
        for a in range(0, n, chunk):
            L.xsi_host_widen_i8_i32(dst.data_ptr() + a, chunk, src.data_ptr() + 4 * a, chunk, np.array([chunk], np.uint32).ctypes.data, 1)
        best = max(best, 4 * n / (time.perf_counter() - t) / 1e9)
    print("widen, calls of %d Mi elements: %.1f GB/s" % (chunk >> 20, best))
a = torch.empty(n, dtype=torch.int32); b = torch.empty(n, dtype=torch.int32)
torch.set_num_threads(16)
for rep in range(3):
    t = time.perf_counter(); b.copy_(a); dt = time.perf_counter() - t
print("torch copy (16 threads): %.1f GB/s read + same written" % (4 * n / dt / 1e9))
