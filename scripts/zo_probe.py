"""ZoomOut ladder 30 -> 200 on a batch of synthetic pairs (cfg4 shape): wall time per pair (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import fm as dfm, nn as dnn, synth, _lib
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 170
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
rng = np.random.default_rng(4000)
n, K = 2000, 200
pool = [synth.synthetic_basis(n, K, rng) for _ in range(4)]
ia, ib = rng.integers(0, 4, P), rng.integers(0, 4, P)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
Phi1, Phi2 = up(np.concatenate([pool[i][1] for i in ia])), up(np.concatenate([pool[i][1] for i in ib]))
a2 = up(np.concatenate([pool[i][2] for i in ib]))
off = np.arange(P + 1) * n
o = dnn.Offsets(torch.from_numpy(off).to(dev), off)
C0 = up(np.stack([np.linalg.qr(rng.standard_normal((30, 30)))[0] for _ in range(P)]))
fl = _lib.DM_FAST_FM if os.environ.get("ZO_FAST") == "1" else 0
fn = lambda: dfm.zoomout(C0, Phi1, Phi2, a2, nit, 1, o, o, return_p2p=True, out_dtype=torch.int32, flags=fl)
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"zoomout 30->{30 + nit}, {P} pairs: {ms:.1f} ms per ladder = {ms / P:.3f} ms per pair = {P / ms * 1e3:.0f} pairs/s")
