"""Timing of the notebook-style fit (w_ent, w_sumto1 on): N = 2000, d = 384, n_ev = 15 / 50."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densematcher_b200.pyFM import FunctionalMapping, TriMesh
from oracle import meshgen
rng = np.random.default_rng(0)
n, d = 2000, 384
for k in (15, 50):
    b1, b2 = meshgen.synthetic_basis(n, k, rng), meshgen.synthetic_basis(n, k, rng)
    coef = rng.standard_normal((k, d))
    c1 = (b1[1] @ coef + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    c2 = (b2[1] @ coef + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    c1 /= np.linalg.norm(c1, axis=1, keepdims=True); c2 /= np.linalg.norm(c2, axis=1, keepdims=True)
    m = FunctionalMapping(TriMesh.from_basis(*b1), TriMesh.from_basis(*b2), optimizer="L-BFGS-B")
    m.preprocess(n_ev=(k, k), descr1=c1, descr2=c2)
    fp = dict(w_descr=1e4, w_lap=1e3, w_dcomm=0, w_ent=1e-1, w_sumto1=1e1, maxiter=5000)
    m.fit(**fp); torch.cuda.synchronize()
    t = time.perf_counter(); m.fit(**fp); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"k={k}: fit with dense terms {dt*1e3:.1f} ms, {m.fit_result.nit} L-BFGS iterations, {m.fit_result.nfev} energy evaluations "
          f"({dt*1e3/m.fit_result.nfev:.2f} ms each)")
    t = time.perf_counter(); m.fit(w_descr=1e4, w_lap=1e3, w_dcomm=0); torch.cuda.synchronize()
    print(f"k={k}: closed-form fit (descr + lap only) {(time.perf_counter()-t)*1e3:.1f} ms")
