"""Spectral ICP with the reference's signatures (densematcher/pyFM/refine/icp.py)."""
from __future__ import annotations

import numpy as np
import torch

from ... import fm as _fm
from .._dev import to_dev


def icp_refine(FM_12, evects1, evects2, A1=None, nit=10, tol=1e-10, use_adj=False, return_p2p=False, n_jobs=1,
               verbose=False):
    """icp.py:43-107.  ``A1`` and ``use_adj`` are accepted for signature parity; as in the reference only
    the p2p_21 search of FM_to_p2p feeds the iteration (icp.py:37).  ``nit`` in (None, 0) iterates until the
    max-abs change of the map is below ``tol`` (icp.py:84-94)."""
    FM_12 = np.asarray(FM_12, dtype=np.float64)
    k2, k1 = FM_12.shape
    evects1, evects2 = np.asarray(evects1), np.asarray(evects2)
    assert k1 <= evects1.shape[1] and k2 <= evects2.shape[1], "At least k eigenvectors should be provided"
    C = to_dev(FM_12, torch.float64)
    P1, P2 = to_dev(evects1[:, :k1], torch.float64), to_dev(evects2[:, :k2], torch.float64)
    if nit is not None and nit > 0:
        res = _fm.icp(C, P1, P2, nit=nit, return_p2p=return_p2p)
    else:
        cur = C[None]
        for _ in range(10000):
            new = _fm.icp(cur, P1, P2, nit=1)
            done = float((new - cur).abs().max()) <= tol
            cur = new
            if done:
                break
        res = (cur, _fm.fm_to_p2p(cur, P1, P2, want=("p2p_21",))["p2p_21"]) if return_p2p else cur
    if return_p2p:
        return res[0][0].cpu().numpy(), res[1].cpu().numpy()
    return res[0].cpu().numpy()


def mesh_icp_refine(FM_12, mesh1, mesh2, nit=10, tol=1e-10, use_adj=False, return_p2p=False, n_jobs=1, verbose=False):
    """icp.py:110-150."""
    k2, k1 = np.asarray(FM_12).shape
    return icp_refine(FM_12, mesh1.eigenvectors[:, :k1], mesh2.eigenvectors[:, :k2], mesh1.A, nit=nit, tol=tol,
                      use_adj=use_adj, return_p2p=return_p2p, n_jobs=n_jobs, verbose=verbose)
