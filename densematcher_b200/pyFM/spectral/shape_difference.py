"""Shape-difference operators of a functional map (densematcher/pyFM/spectral/shape_difference.py:6-52):
k x k host algebra on the map itself, kept for surface parity (``FunctionalMapping.compute_SD``).
``compute_SD`` from a vertex map needs the mesh stiffness matrices, which are outside the hot path."""
from __future__ import annotations

import numpy as np


def area_SD(FM):
    """(k1, k1) area-based shape difference  FM^T FM  (shape_difference.py:6-24)."""
    FM = np.asarray(FM)
    return FM.T @ FM


def conformal_SD(FM, evals1, evals2):
    """(k1, k1) conformal shape difference  pinv(diag(l1)) FM^T diag(l2) FM  (shape_difference.py:27-52)."""
    FM = np.asarray(FM)
    k2, k1 = FM.shape
    return np.linalg.pinv(np.diag(np.asarray(evals1)[:k1])) @ FM.T @ (np.asarray(evals2)[:k2, None] * FM)


def compute_SD(mesh1, mesh2, k1=None, k2=None, p2p=None, SD_type="spectral"):
    raise NotImplementedError("compute_SD needs the stiffness matrices of the meshes; outside the accelerated path "
                              "(SURVEY.md section 2 row 9)")
