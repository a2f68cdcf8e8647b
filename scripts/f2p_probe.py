"""Timing of the FM->p2p stage and its pieces at the bench shape (kernel-only via the phase-skip flags)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import pipeline, fm as dfm, _lib

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
b = bench.make_host_batch(P).to_device(dev)
k = bench.K_EIG
A = dfm.project(b.Phi1, b.area1, b.F1, b.o1, k=k); B = dfm.project(b.Phi2, b.area2, b.F2, b.o2, k=k)
C = dfm.fmap_solve(A, B, b.evals1[:, :k], b.evals2[:, :k], pipeline.fmap_c00(b), bench.W_DESCR, bench.W_LAP)
def tm(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ALL = ("p2p_21", "p2p_12", "dense_21", "dense_12")
f2p = lambda fl=0, want=ALL: dfm.fm_to_p2p(C, b.Phi1[:, :k], b.Phi2[:, :k], b.area1, b.o1, b.o2, want=want, flags=fl, out_dtype=torch.int32)
tag = "PROBE(no epilogue) " if os.environ.get("DM_NN_PROBE") == "1" else ""
print(f"{tag}fm_to_p2p 4 outputs      {tm(f2p):8.3f} ms")
print(f"{tag}fm_to_p2p no recheck     {tm(lambda: f2p(_lib.DM_NO_RECHECK)):8.3f} ms")
for want in (("p2p_21",), ("p2p_12",), ("dense_21",), ("dense_12",), ("p2p_21", "dense_21"), ("p2p_12", "dense_12")):
    print(f"{tag}fm_to_p2p {'+'.join(want):22s} {tm(lambda: f2p(0, want)):8.3f} ms")
