import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import pipeline
P = 128
dev = torch.device("cuda", 0)
host = bench.make_host_batch(P).pin()
kw = dict(k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP, copy=False)
def tm(f, n=5):
    f(); f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
for ch in (8, 12, 16, 24, 32):
    print(f"chunk={ch}: full {tm(lambda: pipeline.match_pairs_host(host, dev, chunk_pairs=ch, **kw)):.2f} ms", flush=True)
# copies only (no compute): the floor of the staged path
orig = pipeline.match_pairs_device
pipeline.match_pairs_device = lambda b, **k: {"nn_p2p_21": torch.zeros(b.F2.shape[0], dtype=torch.int32, device=dev)}
for ch in (16,):
    print(f"chunk={ch}: copies only {tm(lambda: pipeline.match_pairs_host(host, dev, chunk_pairs=ch, **kw)):.2f} ms", flush=True)
pipeline.match_pairs_device = orig
