"""p2p <-> functional-map conversions with the reference's signatures
(densematcher/pyFM/spectral/convert.py)."""
from __future__ import annotations

import numpy as np
import scipy.linalg
import torch

from ... import fm as _fm
from .._dev import to_dev, diag_of, is_diagonal


def p2p_to_FM(p2p_21, evects1, evects2, A2=None):
    """Functional map of a vertex map (convert.py:14-51).

    With a (diagonal) target mass ``A2`` -- 1-D areas, scipy sparse or dense -- the product
    ``evects2.T @ A2 @ evects1[p2p_21]`` runs on the GPU in float64.  Without ``A2`` the reference
    solves a least-squares problem (convert.py:51); here it is solved through the normal equations
    with both Gram products and the Cholesky solve (``dm_spd_solve``) on the GPU.  A (n2, n1) matrix map (convert.py:39) is applied on the host.
    """
    evects1, evects2 = np.asarray(evects1), np.asarray(evects2)
    p = p2p_21
    if getattr(p, "ndim", 1) != 1:  # soft / sparse map: pull back on the host, then the same contraction
        pulled = np.asarray(p @ evects1)
        P1 = to_dev(pulled, torch.float64)
        ident = torch.arange(pulled.shape[0], device=P1.device)
    else:
        P1 = to_dev(evects1, torch.float64)
        ident = to_dev(np.asarray(p, dtype=np.int64))
    P2 = to_dev(evects2, torch.float64)
    if A2 is not None:
        if A2.shape[0] != evects2.shape[0]:
            raise ValueError("Can't compute exact pseudo inverse with subsampled eigenvectors")
        if not is_diagonal(A2):
            raise NotImplementedError("non-diagonal mass matrices are not supported (the reference's are lumped)")
        a2 = to_dev(diag_of(A2, evects2.shape[0]), torch.float64)
        return _fm.p2p_to_fm(ident, P1, P2, a2)[0].cpu().numpy()
    # least squares: (Phi2^T Phi2) C = Phi2^T Phi1[p]
    rhs = _fm.p2p_to_fm(ident, P1, P2, None)[0]
    gram = _fm.p2p_to_fm(torch.arange(P2.shape[0], device=P2.device), P2, P2, None)[0]
    return _fm.spd_solve(gram, rhs)[0].cpu().numpy()


def mesh_p2p_to_FM(p2p_21, mesh1, mesh2, dims=None, subsample=None):
    """convert.py:54-93."""
    if dims is None:
        k1, k2 = len(mesh1.eigenvalues), len(mesh2.eigenvalues)
    elif np.issubdtype(type(dims), np.integer):
        k1 = k2 = dims
    else:
        k1, k2 = dims
    if subsample is None:
        return p2p_to_FM(p2p_21, mesh1.eigenvectors[:, :k1], mesh2.eigenvectors[:, :k2], A2=mesh2.A)
    sub1, sub2 = subsample
    return p2p_to_FM(p2p_21, mesh1.eigenvectors[sub1, :k1], mesh2.eigenvectors[sub2, :k2], A2=None)


def FM_to_p2p(FM_12, evects1, evects2, A1, use_adj=False, n_jobs=1, return_indicator=True):
    """(p2p_21, p2p_12, mapped_indicator) like the reference's modified FM_to_p2p (convert.py:96-147).

    ``use_adj`` is ignored exactly as in the reference (both searches always run, :134-140).
    ``return_indicator=False`` (extension) skips materialising the (n2, n1) float64 matrix, which the
    index outputs never need here.
    """
    FM_12 = np.asarray(FM_12, dtype=np.float64)
    k2, k1 = FM_12.shape
    evects1, evects2 = np.asarray(evects1), np.asarray(evects2)
    assert k1 <= evects1.shape[1], f"At least {k1} should be provided, here only {evects1.shape[1]} are given"
    assert k2 <= evects2.shape[1], f"At least {k2} should be provided, here only {evects2.shape[1]} are given"
    if return_indicator and (evects1.shape[1] != k1 or evects2.shape[1] != k2):
        # convert.py:144 multiplies the UNSLICED bases: same failure as the reference
        raise ValueError(f"matmul: mapped_indicator needs evects pre-sliced to ({k2}, {k1}) columns")
    C = to_dev(FM_12, torch.float64)
    P1, P2 = to_dev(evects1, torch.float64), to_dev(evects2, torch.float64)
    out = _fm.fm_to_p2p(C, P1, P2, None, want=("p2p_21", "p2p_12"))
    p2p_21, p2p_12 = out["p2p_21"].cpu().numpy(), out["p2p_12"].cpu().numpy()
    MI = None
    if return_indicator:
        a1 = to_dev(diag_of(A1, evects1.shape[0]), torch.float64)
        MI = _fm.mapped_indicator(C, P1, P2, a1).cpu().numpy()
    return p2p_21, p2p_12, MI


def mesh_FM_to_p2p(FM_12, mesh1, mesh2, use_adj=False, subsample=None, n_jobs=1):
    """convert.py:149-182."""
    k2, k1 = np.asarray(FM_12).shape
    if subsample is None:
        return FM_to_p2p(FM_12, mesh1.eigenvectors[:, :k1], mesh2.eigenvectors[:, :k2], mesh1.A, use_adj=use_adj,
                         n_jobs=n_jobs)
    sub1, sub2 = subsample
    A1 = diag_of(mesh1.A, mesh1.eigenvectors.shape[0])[sub1]
    return FM_to_p2p(FM_12, mesh1.eigenvectors[sub1, :k1], mesh2.eigenvectors[sub2, :k2], A1, use_adj=use_adj,
                     n_jobs=n_jobs)
