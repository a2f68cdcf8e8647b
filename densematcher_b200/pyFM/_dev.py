"""numpy <-> device plumbing for the reference-facing shims."""
from __future__ import annotations

import numpy as np
import torch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("densematcher_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(a, dtype=None):
    """numpy / torch -> contiguous CUDA tensor (dtype preserved unless given)."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        a = np.asarray(a)
        if a.dtype == np.float16:
            a = a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device(), non_blocking=True).contiguous()


def diag_of(A, n):
    """Vertex areas from whatever the reference passes as a mass matrix: 1-D array, scipy sparse
    diagonal matrix, or dense diagonal matrix."""
    if A is None:
        return None
    if hasattr(A, "diagonal") and getattr(A, "ndim", 2) == 2:
        d = np.asarray(A.diagonal()).ravel()
    else:
        d = np.asarray(A)
        if d.ndim == 2:
            d = np.diag(d)
    if d.shape[0] != n:
        raise ValueError("mass matrix does not match the number of vertices")
    return np.ascontiguousarray(d, dtype=np.float64)


def is_diagonal(A):
    """True if A is 1-D or a (sparse / dense) matrix with no off-diagonal entries."""
    if getattr(A, "ndim", 2) == 1:
        return True
    if hasattr(A, "tocoo"):
        c = A.tocoo()
        return bool(np.all(c.row[c.data != 0] == c.col[c.data != 0]))
    A = np.asarray(A)
    return bool(np.count_nonzero(A - np.diag(np.diag(A))) == 0)
