"""Map-quality metrics with the reference's signatures (densematcher/pyFM/eval/evaluate.py:4-93).
These are O(n) gathers over precomputed geodesic matrices: host numpy, like the reference (SURVEY.md 8f rank 4 --
they close the loop on a dataset run, they are not part of the accelerated path)."""
from __future__ import annotations

import numpy as np


def accuracy(p2p, gt_p2p, D1_geod, return_all=False, sqrt_area=None):
    """Mean geodesic error of a target->source vertex map (evaluate.py:4-36)."""
    dists = np.asarray(D1_geod)[(np.asarray(p2p), np.asarray(gt_p2p))]
    if sqrt_area is not None:
        dists = dists / sqrt_area
    return (dists.mean(), dists) if return_all else dists.mean()


def continuity(p2p, D1_geod, D2_geod, edges):
    """Mean stretch of the target edges under the map (evaluate.py:39-68)."""
    p2p, edges = np.asarray(p2p), np.asarray(edges)
    source_len = np.asarray(D2_geod)[(edges[:, 0], edges[:, 1])]
    target_len = np.asarray(D1_geod)[(p2p[edges[:, 0]], p2p[edges[:, 1]])]
    return np.mean(target_len / source_len)


def coverage(p2p, A):
    """Fraction of the source area hit by the map (evaluate.py:71-93); ``A``: (n1,) areas or an (n1, n1) mass matrix."""
    vert_area = np.asarray(A.sum(1)).flatten() if len(A.shape) == 2 else np.asarray(A)
    return vert_area[np.unique(np.asarray(p2p))].sum() / vert_area.sum()
