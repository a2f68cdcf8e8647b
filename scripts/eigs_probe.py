"""Eigenbasis provider timing: spectral_ops.lbo_eigs (device) vs scipy eigsh shift-invert (the reference's call) on
deformed icospheres.  python scripts/eigs_probe.py [subdivisions] [k]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, torch
from densematcher_b200 import spectral_ops, synth
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 4
k = int(sys.argv[2]) if len(sys.argv) > 2 else 200
V0, F = synth.icosphere(sub)
V = synth.deform(V0, (1.0, 1.3, 0.7))
W, a = synth.cotan_stiffness(V, F), synth.lumped_area(V, F)
for deg in (16, 24, 32, 48):
    spectral_ops.lbo_eigs(W, a, k, degree=deg)
    torch.cuda.synchronize(); t = time.perf_counter()
    ev, Phi, info = spectral_ops.lbo_eigs(W, a, k, degree=deg, return_info=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"lbo_eigs n={len(a)} k={k} degree={deg}: {dt * 1e3:.1f} ms, {info}")
t = time.perf_counter()
wr = spla.eigsh(W.tocsc(), k=k, M=sp.diags(a).tocsc(), sigma=-0.01)[0]
print(f"scipy eigsh shift-invert: {(time.perf_counter() - t) * 1e3:.1f} ms; max |d evals| = {np.abs(np.sort(wr) - ev.cpu().numpy()).max():.2e}")

# dataset throughput: many meshes in flight
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32
Ws, ms = [W] * M, [a] * M
for ns in (1, 8, 16, 32):
    spectral_ops.lbo_eigs_many(Ws[:ns], ms[:ns], k, n_streams=ns)
    torch.cuda.synchronize(); t = time.perf_counter()
    spectral_ops.lbo_eigs_many(Ws, ms, k, n_streams=ns)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"lbo_eigs_many: {M} meshes, {ns} in flight: {dt * 1e3:.0f} ms = {dt / M * 1e3:.1f} ms per mesh")
