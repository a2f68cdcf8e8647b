"""ZoomOut ladder timing (cfg4 shape): N = 2000, k = 30 -> 200, step 1."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densematcher_b200 import fm as dfm, _lib
FLAGS = _lib.DM_FAST_FM if os.environ.get('ZO_FAST') else 0
from oracle import meshgen
P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 170
rng = np.random.default_rng(0)
n, K = 2000, 200
pool = [meshgen.synthetic_basis(n, K, rng) for _ in range(4)]
ia, ib = rng.integers(0, 4, P), rng.integers(0, 4, P)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
Phi1 = dev(np.concatenate([pool[i][1] for i in ia])); Phi2 = dev(np.concatenate([pool[i][1] for i in ib]))
a2 = dev(np.concatenate([pool[i][2] for i in ib]))
off = np.arange(P + 1) * n
C0 = dev(np.stack([np.linalg.qr(rng.standard_normal((30, 30)))[0] for _ in range(P)]))
def run():
    return dfm.zoomout(C0, Phi1, Phi2, a2, nit, 1, off, off, return_p2p=True, out_dtype=torch.int32, flags=FLAGS)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(("fast " if FLAGS else "exact ") + f"zoomout 30->{30+nit} on {P} pairs: {ms:.1f} ms  = {ms/P:.2f} ms/pair, {P/ms*1e3:.0f} pairs/s")
