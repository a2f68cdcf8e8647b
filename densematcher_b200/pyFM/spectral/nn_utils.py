"""``knn_query`` with the reference's signature (densematcher/pyFM/spectral/nn_utils.py:4-38)."""
from __future__ import annotations

import numpy as np
import torch

from ... import nn as _nn
from .._dev import to_dev


def knn_query(X, Y, k=1, return_distance=False, n_jobs=1):
    """Nearest neighbour of each row of ``Y`` among the rows of ``X`` (Euclidean).

    Same contract as the reference (a kd-tree there, nn_utils.py:28-30): returns ``matches``
    (n2,) int64, or ``(dists, matches)`` with float64 distances.  ``n_jobs`` is accepted and
    ignored (the search is one fused GPU pass).  ``k == 1`` is the hot path; ``k > 1`` (used by the reference's
    barycentric "precise map", projection_utils.py:178) returns (n2, k) arrays through ``dm_knn_f64`` (k <= 16).
    float32 inputs are scored as they are, any other dtype as float64 (sklearn casts to float64).
    """
    X = np.asarray(X) if not isinstance(X, torch.Tensor) else X
    Y = np.asarray(Y) if not isinstance(Y, torch.Tensor) else Y
    if X.ndim != 2 or Y.ndim != 2 or X.shape[1] != Y.shape[1]:
        raise ValueError(f"X {tuple(X.shape)} and Y {tuple(Y.shape)} must be 2-D with the same width")
    if X.shape[0] == 0:
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required")  # sklearn's message
    if k != 1:
        # the k > 1 form (projection_utils.py:178): (n2, k) arrays ordered by distance, float64 like sklearn
        if k > X.shape[0]:
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = {X.shape[0]}")
        if Y.shape[0] == 0:
            e = np.zeros((0, k))
            return (e, e.astype(np.int64)) if return_distance else e.astype(np.int64)
        d, m = _nn.knn_topk(to_dev(Y, torch.float64), to_dev(X, torch.float64), int(k))
        return (d.cpu().numpy(), m.cpu().numpy()) if return_distance else m.cpu().numpy()
    both32 = str(X.dtype).endswith("float32") and str(Y.dtype).endswith("float32")
    dt = torch.float32 if both32 else torch.float64
    Xd, Yd = to_dev(X, dt), to_dev(Y, dt)
    (idx,), _ = _nn.nn_argmax(Yd, Xd, row_epi=(_nn.EUCLID,))
    matches = idx.cpu().numpy()
    if not return_distance:
        return matches
    if both32:
        d = _nn.match_dist(Yd, Xd, idx).cpu().numpy()
    else:
        d = torch.linalg.vector_norm(Yd - Xd[idx], dim=1).cpu().numpy()
    return d, matches
