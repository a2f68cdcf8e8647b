"""ZoomOut with the reference's signatures (densematcher/pyFM/refine/zoomout.py) and upstream pyFM
semantics: the shipped reference calls a stale ``FM_to_p2p`` signature and raises TypeError
(zoomout.py:40,112 vs convert.py:96), so the loop follows zoomout.py:7-115 with the
un-modified conversion it was written against."""
from __future__ import annotations

import numpy as np
import torch

from ... import fm as _fm
from .._dev import to_dev, diag_of, is_diagonal
from ..spectral import convert as _convert
from ..spectral.nn_utils import knn_query


def _steps(step):
    try:
        s1, s2 = step
    except TypeError:
        s1 = s2 = step
    return int(s1), int(s2)


def zoomout_iteration(FM_12, evects1, evects2, step=1, A2=None, n_jobs=1):
    """One step (zoomout.py:7-44)."""
    return zoomout_refine(FM_12, evects1, evects2, nit=1, step=step, A2=A2)


def zoomout_refine(FM_12, evects1, evects2, nit=10, step=1, A2=None, subsample=None, return_p2p=False, n_jobs=1,
                   verbose=False):
    """zoomout.py:47-115.  ``A2``: target vertex areas (1-D, sparse or dense diagonal).  With ``subsample``
    (a pair of index arrays) the map is fitted by least squares on the samples (:104-105)."""
    FM_12 = np.asarray(FM_12, dtype=np.float64)
    evects1, evects2 = np.asarray(evects1), np.asarray(evects2)
    k2_0, k1_0 = FM_12.shape
    s1, s2 = _steps(step)
    assert k1_0 + nit * s1 <= evects1.shape[1], \
        f"Not enough eigenvectors on source : {k1_0 + nit*s1} are needed when {evects1.shape[1]} are provided"
    assert k2_0 + nit * s2 <= evects2.shape[1], \
        f"Not enough eigenvectors on target : {k2_0 + nit*s2} are needed when {evects2.shape[1]} are provided"
    use_sub = subsample is not None
    if use_sub or A2 is None:
        # least-squares variant: host loop over GPU primitives (not the batched ladder kernel path)
        E1 = evects1[subsample[0]] if use_sub else evects1
        E2 = evects2[subsample[1]] if use_sub else evects2
        C = FM_12
        for _ in range(nit):
            k2, k1 = C.shape
            p = knn_query(E1[:, :k1] @ C.T, E2[:, :k2])
            C = _convert.p2p_to_FM(p, E1[:, :k1 + s1], E2[:, :k2 + s2], A2=None)
        if return_p2p:
            k2, k1 = C.shape
            return C, knn_query(evects1[:, :k1] @ C.T, evects2[:, :k2])
        return C
    if A2.shape[0] != evects2.shape[0]:
        raise ValueError("Can't compute exact pseudo inverse with subsampled eigenvectors")
    if not is_diagonal(A2):
        raise NotImplementedError("non-diagonal mass matrices are not supported")
    a2 = to_dev(diag_of(A2, evects2.shape[0]), torch.float64)
    P1 = to_dev(evects1[:, :k1_0 + nit * s1], torch.float64)
    P2 = to_dev(evects2[:, :k2_0 + nit * s2], torch.float64)
    res = _fm.zoomout(to_dev(FM_12, torch.float64), P1, P2, a2, nit, (s1, s2), return_p2p=return_p2p)
    if return_p2p:
        return res[0][0].cpu().numpy(), res[1].cpu().numpy()
    return res[0].cpu().numpy()


def mesh_zoomout_refine(FM_12, mesh1, mesh2, nit=10, step=1, subsample=None, return_p2p=False, n_jobs=1,
                        verbose=False):
    """zoomout.py:118-161.  An integer ``subsample`` draws a farthest point sample of that size on both meshes
    (:150-156) -- Euclidean, on the GPU (``TriMesh.extract_fps``; the reference's geodesic variant needs the un-vendored
    potpourri3d heat method)."""
    if np.issubdtype(type(subsample), np.integer):
        subsample = (mesh1.extract_fps(subsample, geodesic=False), mesh2.extract_fps(subsample, geodesic=False))
    return zoomout_refine(FM_12, mesh1.eigenvectors, mesh2.eigenvectors, nit, step=step, A2=mesh2.A,
                          subsample=subsample, return_p2p=return_p2p, n_jobs=n_jobs, verbose=verbose)


def mesh_zoomout_refine_p2p(p2p_21, mesh1, mesh2, k_init, nit=10, step=1, subsample=None, return_p2p=False, n_jobs=1,
                            p2p_on_sub=False, verbose=False):
    """zoomout.py:164-217: start the ladder from a vertex map."""
    if np.issubdtype(type(subsample), np.integer):
        if p2p_on_sub:
            raise ValueError("P2P can't be defined on undefined subsample")          # zoomout.py:200-201
        subsample = (mesh1.extract_fps(subsample, geodesic=False), mesh2.extract_fps(subsample, geodesic=False))
    k1_0, k2_0 = (k_init, k_init) if np.issubdtype(type(k_init), np.integer) else k_init
    if subsample is None or not p2p_on_sub:
        # the vertex map lives on the full meshes: initial map from all the vertices (zoomout.py:211 ->
        # convert.py:89-90), the ladder itself on the samples
        FM_12 = _convert.p2p_to_FM(p2p_21, mesh1.eigenvectors[:, :k1_0], mesh2.eigenvectors[:, :k2_0], A2=mesh2.A)
    else:
        sub1, sub2 = subsample
        FM_12 = _convert.p2p_to_FM(p2p_21, mesh1.eigenvectors[sub1, :k1_0], mesh2.eigenvectors[sub2, :k2_0], A2=None)
    return mesh_zoomout_refine(FM_12, mesh1, mesh2, nit=nit, step=step, subsample=subsample, return_p2p=return_p2p,
                               n_jobs=n_jobs, verbose=verbose)
