"""Where does the host-buffer entry spend its time?  (run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import pipeline

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
t0 = time.perf_counter(); host = bench.make_host_batch(P); t1 = time.perf_counter(); host.pin(); t2 = time.perf_counter()
print(f"make {t1-t0:.2f}s pin {t2-t1:.2f}s bytes {host.h2d_bytes()/1e6:.0f} MB", flush=True)
def tm(f, n=5):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
big = host._pinned["F1"]
print("pinned?", big.is_pinned(), big[: big.shape[0] // 2].is_pinned())
dst = torch.empty_like(big, device=dev)
ms = tm(lambda: dst.copy_(big, non_blocking=True)); print(f"H2D F1 {big.nbytes/1e6:.0f} MB: {ms:.2f} ms = {big.nbytes/ms/1e6:.1f} GB/s", flush=True)
ms = tm(lambda: host.to_device(dev)); print(f"to_device whole batch: {ms:.2f} ms = {host.h2d_bytes()/ms/1e6:.1f} GB/s", flush=True)
d = host.to_device(dev)
kw = dict(k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP)
ms = tm(lambda: pipeline.match_pairs_device(d, **kw)); print(f"device pipeline: {ms:.2f} ms", flush=True)
res = pipeline.match_pairs_device(d, **kw)
def d2h():
    return {n: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t, non_blocking=True) for n, t in res.items()}
ms = tm(d2h); print(f"D2H (fresh pinned): {ms:.2f} ms", flush=True)
for ch in (8, 16, 32, 64, 128):
    ms = tm(lambda: pipeline.match_pairs_host(host, dev, chunk_pairs=ch, **kw), 3); print(f"match_pairs_host chunk={ch}: {ms:.2f} ms -> {P/ms*1e3:.0f} pairs/s", flush=True)
