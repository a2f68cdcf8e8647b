"""Drop-in ``compute_surface_map`` (densematcher/functional_map.py:9-81): same signature, same 14-tuple.

    (p2p_21, p2p_12, hungarian, hungarian_precise, p2p_21_icp, p2p_12_icp, hungarian_icp, model, model.mesh1,
     model.mesh2, p2p_21_adjoint, p2p_12_adjoint, p2p_21_icp_adjoint, p2p_12_icp_adjoint)

Slots 0/1 and 4/5 are the dense-argmax maps (functional_map.py:49-50, :76-77), the ``*_adjoint`` slots the
kd-tree-equivalent searches of FM_to_p2p (:48, :75) -- all four come out of one fused GPU pass per map.
The Hungarian assignments (:57, :66, :78) are solved in HBM by ``dm_lap_solve`` -- the same assignment scipy returns,
tie-breaking included (``hungarian_icp`` always, like the reference; ``hungarian=False`` skips them,
``hungarian="scipy"`` runs scipy on the host instead); with ``compute_extra`` the barycentric precise map (:62,
``dm_precise_map``) feeds ``hungarian_precise``.

``mesh1_t`` / ``mesh2_t`` may be pytorch3d-like objects (only ``verts_list()[0]`` / ``faces_list()[0]`` are read,
:17-18) or ``densematcher_b200.pyFM.mesh.TriMesh`` instances that already carry a spectrum (the accelerated
path's input contract: precomputed eigenbases).
"""
from __future__ import annotations

import numpy as np

from .pyFM.functional import FunctionalMapping
from .pyFM.mesh import TriMesh

__all__ = ["compute_surface_map"]


def _as_trimesh(m):
    if isinstance(m, TriMesh):
        return m
    if hasattr(m, "eigenvectors") and hasattr(m, "A"):      # a reference TriMesh (duck-typed)
        t = TriMesh(getattr(m, "vertlist", None), getattr(m, "facelist", None))
        t.eigenvalues, t.eigenvectors, t.A = m.eigenvalues, m.eigenvectors, m.A
        return t
    return TriMesh(m.verts_list()[0].cpu(), m.faces_list()[0].cpu())


def compute_surface_map(mesh1_t, mesh2_t, c1, c2, n_ev=50, compute_extra=False, optimizer="fmin_l_bfgs_b",
                        descr_type="neural", maxiter=100000, optimize_p2p=False, fit_params=None, hungarian=True):
    assert descr_type in ["neural", "HKS", "WKS"]
    if descr_type != "neural":
        raise NotImplementedError("HKS / WKS descriptors are outside the hot path (SURVEY.md section 2 row 10)")
    from scipy.optimize import linear_sum_assignment

    def to_np(c):
        return c.detach().cpu().numpy() if hasattr(c, "detach") else np.asarray(c)

    mesh1, mesh2 = _as_trimesh(mesh1_t), _as_trimesh(mesh2_t)
    model = FunctionalMapping(mesh1, mesh2, partial=False, optimizer=optimizer)
    model.preprocess(n_ev=(n_ev, n_ev), n_descr=c1.shape[1], landmarks=None, descr1=to_np(c1), descr2=to_np(c2),
                     subsample_step=1)
    model.fit(**fit_params)  # fit_params=None raises TypeError like the reference (functional_map.py:47)

    def assign():
        if hungarian == "scipy":                              # host solver on the materialised indicator
            mi = model.mapped_indicator * model.eta[..., None] - 1000 * (1 - model.eta[..., None])
            return linear_sum_assignment(mi, maximize=True)
        return model.hungarian()                              # dm_lap_solve: identical assignment, in HBM

    p2p_21_adjoint, p2p_12_adjoint, p2p_21, p2p_12 = model.get_p2p(n_jobs=1, dense=True)
    hung = assign() if (compute_extra and hungarian) else None
    hung_precise = None
    if compute_extra and hungarian:                           # functional_map.py:60-66
        precise = model.get_precise_map().toarray()
        if hungarian == "scipy":
            eta = model.eta[..., None]
            hung_precise = linear_sum_assignment(precise * eta - 1000 * (1 - eta), maximize=True)
        else:
            hung_precise = model.hungarian(indicator=precise)
    model.icp_refine()
    p2p_21_icp_adjoint, p2p_12_icp_adjoint, p2p_21_icp, p2p_12_icp = model.get_p2p(n_jobs=1, dense=True)
    hung_icp = assign() if hungarian else None
    return (p2p_21, p2p_12, hung, hung_precise, p2p_21_icp, p2p_12_icp, hung_icp, model, model.mesh1, model.mesh2,
            p2p_21_adjoint, p2p_12_adjoint, p2p_21_icp_adjoint, p2p_12_icp_adjoint)
