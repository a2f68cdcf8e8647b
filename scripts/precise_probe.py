"""Times dm_precise_map (icosphere(4) pair, 2562 vertices / 5120 faces, p = 50) alone and in a batch of 32 pairs."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import meshgen
from densematcher_b200 import fm

V, F = meshgen.icosphere(4)
V1 = meshgen.deform(V, (1.0, 0.9, 1.1), 0.1, (0.3, 0.2))
V2 = meshgen.deform(V, (1.1, 1.0, 0.85), 0.15, (1.0, 0.5))
k = 50
ev1, P1, a1 = meshgen.lbo_basis(V1, F, k)
ev2, P2, a2 = meshgen.lbo_basis(V2, F, k)
C = P2.T @ (a2[:, None] * P1)          # the functional map of the identity vertex map
emb1, emb2 = P1, P2 @ C
n, nf = len(V), len(F)
for nb in (1, 32):
    e1 = torch.from_numpy(np.tile(emb1, (nb, 1))).cuda()
    e2 = torch.from_numpy(np.tile(emb2, (nb, 1))).cuda()
    ff = torch.from_numpy(np.tile(F.astype(np.int32), (nb, 1))).cuda()
    off = np.arange(nb + 1) * n
    foff = np.arange(nb + 1) * nf
    fm.precise_map(e1, ff, e2, off, foff, off)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(3):
        face, bary = fm.precise_map(e1, ff, e2, off, foff, off)
    torch.cuda.synchronize(); dt = (time.time() - t) / 3
    print(f"precise map, batch {nb:3d}: {dt*1e3:8.2f} ms total, {dt*1e3/nb:7.2f} ms / pair")
b = bary.cpu().numpy()
print("bary rows sum to 1:", np.abs(b.sum(1) - 1).max(), " vertex-hit fraction:", float((b.max(1) > 1 - 1e-9).mean()))
