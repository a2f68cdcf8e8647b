import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densematcher_b200 import fm as dfm
from oracle import meshgen
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0); n, K = 2000, 100
pool = [meshgen.synthetic_basis(n, K, rng) for _ in range(4)]
ia, ib = rng.integers(0, 4, P), rng.integers(0, 4, P)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
Phi1 = dev(np.concatenate([pool[i][1] for i in ia])); Phi2 = dev(np.concatenate([pool[i][1] for i in ib]))
off = np.arange(P + 1) * n
C0 = dev(np.stack([np.linalg.qr(rng.standard_normal((K, K)))[0] for _ in range(P)]))
run = lambda: dfm.icp(C0, Phi1, Phi2, 10, off, off, return_p2p=True, out_dtype=torch.int32)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"icp nit=10 k={K} on {P} pairs: {e0.elapsed_time(e1):.1f} ms")
