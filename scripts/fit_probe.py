"""Throughput of the batched on-device fit with the notebook's dense-map terms (N = 2000, k = 30, d = 384)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import fm as dfm, pipeline
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
k = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
b = bench.make_host_batch(P).to_device(dev)
A = dfm.project(b.Phi1, b.area1, b.F1, b.o1, k=k); B = dfm.project(b.Phi2, b.area2, b.F2, b.o2, k=k)
c00 = pipeline.fmap_c00(b)
Phi1, Phi2 = b.Phi1[:, :k].contiguous(), b.Phi2[:, :k].contiguous()
w = {"ent": 1e-1, "sumto1": 1e1}
call = lambda: dfm.fit_dense(A, B, b.evals1[:, :k], b.evals2[:, :k], c00, Phi1, Phi2, b.area1, w, 1e4, 1e3, off1=b.o1, off2=b.o2, return_info=True)
C, info = call(); torch.cuda.synchronize()
t0 = time.perf_counter(); C, info = call(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dfm.dense_energy(C, Phi1, Phi2, b.area1, w, b.o1, b.o2); e1.record(); torch.cuda.synchronize()
print(f"fit_dense: {P} pairs, k = {k}: {dt * 1e3:.1f} ms = {P / dt:.0f} pairs/s; iterations {info[0]}, energy evaluations {info[1]}; one dm_dense_energy launch {e0.elapsed_time(e1):.2f} ms")

if "--real" in sys.argv:
    # a real LBO pair (two deformations of icosphere(4), 2562 vertices) with band-limited descriptors, replicated P times
    from densematcher_b200 import synth
    V, F = synth.icosphere(4)
    ev1, Q1, a1 = synth.lbo_basis(synth.deform(V, (1.0, 1.3, 0.7)), F, k)
    ev2, Q2, a2 = synth.lbo_basis(synth.deform(V, (1.2, 0.8, 1.0), bump=0.15, phase=(0.3, 1.1)), F, k)
    rng = np.random.default_rng(1)
    coef = rng.standard_normal((min(k, 30), 64))
    c1 = Q1[:, :coef.shape[0]] @ coef + 0.02 * rng.standard_normal((len(V), 64))
    c2 = Q2[:, :coef.shape[0]] @ coef + 0.02 * rng.standard_normal((len(V), 64))
    c1 = (c1 / np.linalg.norm(c1, axis=1, keepdims=True)).astype(np.float32)
    c2 = (c2 / np.linalg.norm(c2, axis=1, keepdims=True)).astype(np.float32)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    n = len(V)
    off = np.arange(P + 1) * n
    rep = lambda a: up(np.concatenate([a] * P))
    Phi1, Phi2, ar1, ar2 = rep(Q1), rep(Q2), rep(a1), rep(a2)
    A = dfm.project(Phi1, ar1, rep(c1), off, k=k); B = dfm.project(Phi2, ar2, rep(c2), off, k=k)
    sg = np.sign(Q1[0, 0] * Q2[0, 0]) * np.sqrt(a2.sum() / a1.sum())
    call = lambda: dfm.fit_dense(A, B, up(np.tile(ev1, (P, 1))), up(np.tile(ev2, (P, 1))), up(np.full(P, sg)), Phi1, Phi2, ar1,
                                 w, 1e4, 1e3, off1=off, off2=off, return_info=True)
    C, info = call(); torch.cuda.synchronize()
    t0 = time.perf_counter(); C, info = call(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"fit_dense (real LBO pair, N = {n}): {P} pairs, k = {k}: {dt * 1e3:.1f} ms = {P / dt:.0f} pairs/s; iterations {info[0]}, "
          f"energy evaluations {info[1]}")
