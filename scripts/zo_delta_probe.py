"""How many entries of p2p_21 change from one ZoomOut rung to the next (decides whether an incremental p2p -> FM update
would pay): bench-like synthetic pairs and the full-size reference golden."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import fm as dfm, nn as dnn, synth
dev = torch.device("cuda", 0)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

def ladder(C0, Phi1, Phi2, a2, k0, nit, tag):
    C = C0
    prev = None
    out = []
    for it in range(nit):
        C, p = dfm.zoomout(C, Phi1, Phi2, a2, 1, 1, return_p2p=False), None
        # p2p of the NEW C is what the next rung uses; ask for it explicitly
        _, p = dfm.zoomout(C, Phi1, Phi2, a2, 0, 1, return_p2p=True)
        if prev is not None:
            out.append(float((p != prev).float().mean()))
        prev = p
    print(tag, "changed fraction per rung (every 10th):", " ".join(f"{x:.3f}" for x in out[::10]), "mean %.3f" % np.mean(out))

g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fm_full_ico4.npz")))
P1, P2 = up(g["Phi1_f32"].astype(np.float64)), up(g["Phi2_f32"].astype(np.float64))
C0 = up(g["C_closed_form"][:30, :30])[None]
ladder(C0, P1, P2, up(g["area2"]), 30, 170, "golden ico4:")
rng = np.random.default_rng(4000)
n, K = 2000, 200
pool = [synth.synthetic_basis(n, K, rng) for _ in range(2)]
Phi1, Phi2, a2 = up(pool[0][1]), up(pool[1][1]), up(pool[1][2])
C0 = up(np.linalg.qr(rng.standard_normal((30, 30)))[0])[None]
ladder(C0, Phi1, Phi2, a2, 30, 170, "synthetic QR bases:")
