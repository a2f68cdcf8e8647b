"""tcgen05 projection vs float64: error and time (run on the GPU box under `timeout`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densematcher_b200 import fm, _lib
from oracle import meshgen, dm_oracle as orc
rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
relF = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
for (n, k, d, nb) in [(64, 8, 64, 1), (300, 20, 32, 1), (2000, 100, 384, 1), (2000, 100, 384, 3), (1777, 130, 200, 2)]:
    ns = [n - 37 * i for i in range(nb)]
    bases = [meshgen.synthetic_basis(m, k, rng) for m in ns]
    Phi = np.concatenate([b[1] for b in bases]); area = np.concatenate([b[2] for b in bases])
    F = meshgen.random_unit_features(sum(ns), d, rng)
    off = np.concatenate([[0], np.cumsum(ns)])
    ref = np.stack([orc.project(Phi[off[i]:off[i+1]], area[off[i]:off[i+1]], F[off[i]:off[i+1]], k) for i in range(nb)])
    t0 = time.time()
    out64 = fm.project(dev(Phi), dev(area), dev(F), off, k=k, flags=_lib.DM_F64_GEMM).cpu().numpy()
    print(f"n={ns} k={k} d={d}: f64 relF {relF(out64, ref):.2e}", flush=True)
    out = fm.project(dev(Phi), dev(area), dev(F), off, k=k).cpu().numpy()
    e = np.abs(out - ref)
    print(f"   tc relF {relF(out, ref):.2e}  max abs err {e.max():.2e} (ref max {np.abs(ref).max():.2e})  per-batch relF {[float('%.1e' % relF(out[i], ref[i])) for i in range(nb)]}  ({time.time()-t0:.2f}s)", flush=True)
    if relF(out, ref) > 1e-3:
        i = 0
        bad = np.argwhere(e[i] > 1e-3 * np.abs(ref[i]).max())
        print("   BAD entries", len(bad), "rows", np.unique(bad[:, 0])[:20], "cols", np.unique(bad[:, 1])[:20], flush=True)
        print("   out[0,:4,:6]\n", out[0, :4, :6], "\n   ref\n", ref[0, :4, :6], flush=True)
print("probe done")
