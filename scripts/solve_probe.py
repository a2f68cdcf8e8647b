"""A/B timing of the closed-form solve kernels at the bench shape (128 pairs, k = 100, d = 384)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import pipeline, fm as dfm

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
b = bench.make_host_batch(P).to_device(dev)
k = bench.K_EIG
A = dfm.project(b.Phi1, b.area1, b.F1, b.o1, k=k); B = dfm.project(b.Phi2, b.area2, b.F2, b.o2, k=k)
c00 = pipeline.fmap_c00(b)
def tm(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
call = lambda **kw: dfm.fmap_solve(A, B, b.evals1[:, :k], b.evals2[:, :k], c00, bench.W_DESCR, bench.W_LAP, **kw)
ref = None
for mode in ("f64", "", "f32t64"):
    if mode: os.environ["DM_SOLVE"] = mode
    else: os.environ.pop("DM_SOLVE", None)
    C, st = call(return_status=True)
    if ref is None: ref = C
    rel = float((C - ref).norm() / ref.norm())
    print(f"solve mode {mode or 'f32 (default)':16s} {tm(lambda: call(check=False)):8.3f} ms   status {st}   relF vs f64 {rel:.2e}", flush=True)
os.environ.pop("DM_SOLVE", None)
