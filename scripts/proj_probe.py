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
# effect on C (N=2000, k=100, d=384, notebook weights)
n, k, d = 2000, 100, 384
b1, b2 = meshgen.synthetic_basis(n, k, rng), meshgen.synthetic_basis(n, k, rng)
F1, F2 = meshgen.random_unit_features(n, d, rng), meshgen.random_unit_features(n, d, rng)
c00 = orc.fmap_c00(b1[1], b2[1], b1[2], b2[2])
Cs = {}
for name, fl in (("f64", _lib.DM_F64_GEMM), ("tc", 0)):
    A = fm.project(dev(b1[1]), dev(b1[2]), dev(F1), k=k, flags=fl); B = fm.project(dev(b2[1]), dev(b2[2]), dev(F2), k=k, flags=fl)
    Cs[name] = fm.fmap_solve(A, B, dev(b1[0])[None], dev(b2[0])[None], dev(np.array([c00])), 1e4, 1e3)[0].cpu().numpy()
Co = orc.fmap_solve_closed_form(orc.project(b1[1], b1[2], F1, k), orc.project(b2[1], b2[2], F2, k), b1[0], b2[0], c00, 1e4, 1e3)
print(f"C relF vs oracle closed form: f64 projection {relF(Cs['f64'], Co):.2e}, tc projection {relF(Cs['tc'], Co):.2e}")
