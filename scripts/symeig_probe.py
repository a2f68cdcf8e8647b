"""Dense symmetric eigensolver timing (one CTA per matrix): python scripts/symeig_probe.py [m] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densematcher_b200 import spectral_ops
from densematcher_b200.nn import default_workspace
m = int(sys.argv[1]) if len(sys.argv) > 1 else 240
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
A = rng.standard_normal((B, m, m)); A = A + A.transpose(0, 2, 1)
Ad = torch.from_numpy(A).cuda()
for kind in ("random", "near-identity"):
    if kind == "near-identity":
        Ad = torch.eye(m, dtype=torch.float64, device="cuda")[None] + 1e-3 * Ad
    spectral_ops.sym_eig(Ad); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w, V = spectral_ops.sym_eig(Ad); e1.record(); torch.cuda.synchronize()
    st = default_workspace(Ad.device, "eig").buf[:256].view(torch.int32).cpu().numpy()
    print(f"{kind}: sym_eig m={m} batch={B}: {e0.elapsed_time(e1):.2f} ms; kcycles tridiag={st[8]} formQ={st[9]} ql={st[10]}")
