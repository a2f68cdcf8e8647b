"""Times dm_lap_solve on mapped-indicator matrices (icosphere(4) pair, 2562 x 2562) against scipy on the host."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.optimize import linear_sum_assignment
from densematcher_b200 import fm
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_lap import _mapped_indicator

mi = _mapped_indicator(4, 50, 1)
t = time.time(); rr, cc = linear_sum_assignment(mi, maximize=True); t_scipy = time.time() - t
print(f"scipy (1 core)          {t_scipy*1e3:9.1f} ms / problem")
d = torch.from_numpy(mi).cuda()
for nb in (1, 16, 148, 296):
    rng = np.random.default_rng(nb)
    mats = [d] + [d + 1e-7 * torch.from_numpy(rng.standard_normal(mi.shape)).cuda() for _ in range(min(nb, 8) - 1)]
    mats = [mats[i % len(mats)] for i in range(nb)]
    fm.lap_solve(mats, maximize=True)
    torch.cuda.synchronize(); t = time.time()
    res = fm.lap_solve(mats, maximize=True)
    torch.cuda.synchronize(); dt = time.time() - t
    ok = np.array_equal(res[0][1], cc)
    print(f"gpu batch {nb:4d}          {dt*1e3:9.1f} ms total  {dt*1e3/nb:9.2f} ms / problem   identical to scipy: {ok}")
