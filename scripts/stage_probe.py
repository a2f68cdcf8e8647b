"""Per-stage device timings of one bench step (CUDA events, warm), plus the re-evaluation statistics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import pipeline, nn as dnn, fm as dfm, _lib

P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
host = bench.make_host_batch(P)
b = host.to_device(dev)
k = bench.K_EIG
def tm(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nnc = lambda fl=0, **kw: dnn.nn_argmax(b.F2, b.F1, b.off2, b.off1, row_epi=(dnn.COSINE_UNIT,), col_epi=(dnn.COSINE_UNIT,),
                                 max_q=b.max2, max_db=b.max1, flags=fl, out_dtype=torch.int32, **kw)
print("pairs", P)
print("feature NN stats (rows, cols, full scans):", nnc(return_stats=True)[2])
print(f"feature NN stage        {tm(nnc):8.3f} ms")
print(f"  score kernel only     {tm(lambda: nnc(_lib.DM_SKIP_PREP | _lib.DM_SKIP_FINISH)):8.3f} ms")
print(f"  no recheck            {tm(lambda: nnc(_lib.DM_NO_RECHECK)):8.3f} ms")
print(f"  kernel+finish         {tm(lambda: nnc(_lib.DM_SKIP_PREP)):8.3f} ms")
rowonly = lambda fl: dnn.nn_argmax(b.F2, b.F1, b.off2, b.off1, row_epi=(dnn.COSINE_UNIT,), max_q=b.max2, max_db=b.max1, flags=fl, out_dtype=torch.int32)
colonly = lambda fl: dnn.nn_argmax(b.F2, b.F1, b.off2, b.off1, row_epi=(), col_epi=(dnn.COSINE_UNIT,), max_q=b.max2, max_db=b.max1, flags=fl, out_dtype=torch.int32)
rowonly(0); colonly(0)
print(f"  score kernel row-only {tm(lambda: rowonly(_lib.DM_SKIP_PREP | _lib.DM_SKIP_FINISH)):8.3f} ms")
print(f"  score kernel col-only {tm(lambda: colonly(_lib.DM_SKIP_PREP | _lib.DM_SKIP_FINISH)):8.3f} ms")
A = dfm.project(b.Phi1, b.area1, b.F1, b.o1, k=k); B = dfm.project(b.Phi2, b.area2, b.F2, b.o2, k=k)
print(f"project (one mesh side) {tm(lambda: dfm.project(b.Phi1, b.area1, b.F1, b.o1, k=k)):8.3f} ms")
c00 = pipeline.fmap_c00(b)
print(f"c00 (torch ops)         {tm(lambda: pipeline.fmap_c00(b)):8.3f} ms")
C = dfm.fmap_solve(A, B, b.evals1[:, :k], b.evals2[:, :k], c00, bench.W_DESCR, bench.W_LAP)
print(f"fmap_solve              {tm(lambda: dfm.fmap_solve(A, B, b.evals1[:, :k], b.evals2[:, :k], c00, bench.W_DESCR, bench.W_LAP)):8.3f} ms")
f2p = lambda fl=0, want=("p2p_21", "p2p_12", "dense_21", "dense_12"): dfm.fm_to_p2p(C, b.Phi1[:, :k], b.Phi2[:, :k], b.area1, b.o1, b.o2, want=want, flags=fl, out_dtype=torch.int32)
print(f"fm_to_p2p (4 outputs)   {tm(f2p):8.3f} ms")
print(f"fm_to_p2p no recheck    {tm(lambda: f2p(_lib.DM_NO_RECHECK)):8.3f} ms")
print(f"fm_to_p2p p2p_21 only   {tm(lambda: f2p(0, ('p2p_21',))):8.3f} ms")
for want in (("p2p_21", "dense_21"), ("p2p_12", "dense_12"), ("p2p_12",), ("dense_12",), ("p2p_21", "p2p_12")):
    print(f"fm_to_p2p {'+'.join(want):22s} {tm(lambda: f2p(0, want)):8.3f} ms")
print(f"whole step              {tm(lambda: pipeline.match_pairs_device(b, k=k, w_descr=bench.W_DESCR, w_lap=bench.W_LAP)):8.3f} ms")
if "--icp" in sys.argv:
    print(f"icp nit=10              {tm(lambda: dfm.icp(C, b.Phi1[:, :k], b.Phi2[:, :k], 10, b.o1, b.o2), 2):8.3f} ms")
    print(f"zoomout 30->50          {tm(lambda: dfm.zoomout(C[:, :30, :30].contiguous(), b.Phi1, b.Phi2, b.area2, 20, 1, b.o1, b.o2), 2):8.3f} ms")
