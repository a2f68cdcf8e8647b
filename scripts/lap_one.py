"""One dm_lap_solve call on a 2562 x 2562 mapped indicator (profiling target)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from densematcher_b200 import fm
from test_gpu_lap import _mapped_indicator
mi = torch.from_numpy(_mapped_indicator(4, 50, 1)).cuda()
fm.lap_solve([mi], maximize=True)
torch.cuda.synchronize()
