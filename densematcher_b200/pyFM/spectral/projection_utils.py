"""Barycentric precise map with the reference's signatures (densematcher/pyFM/spectral/projection_utils.py):
``project_pc_to_triangles`` and ``barycentric_to_precise``.  The projection runs in ``dm_precise_map``."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sparse
import torch

from ... import fm as _fm
from .._dev import to_dev


def barycentric_to_precise(faces, face_match, bary_coord, n_vertices=None):
    """(n2, n1) sparse map from face indices and barycentric coordinates (projection_utils.py:380-417)."""
    faces = np.asarray(faces)
    if n_vertices is None:
        n_vertices = 1 + faces.max()
    n_points = face_match.shape[0]
    I = np.tile(np.arange(n_points), 3)
    J = np.concatenate([faces[face_match, 0], faces[face_match, 1], faces[face_match, 2]])
    S = np.concatenate([bary_coord[:, 0], bary_coord[:, 1], bary_coord[:, 2]])
    return sparse.csr_matrix((S, (I, J)), shape=(n_points, n_vertices))


def project_pc_to_triangles(vert_emb, faces, points_emb, precompute_dmin=True, batch_size=None, n_jobs=1, verbose=False,
                            return_bary=False):
    """Projects a p-dimensional point cloud on a p-dimensional triangle mesh (projection_utils.py:16-115).
    ``precompute_dmin`` / ``batch_size`` / ``n_jobs`` only chose between equivalent host strategies in the reference
    and are accepted for compatibility.  ``return_bary=True`` additionally returns (face_match, bary_coord)."""
    vert_emb, points_emb = np.asarray(vert_emb), np.asarray(points_emb)
    fm_d, bary_d = _fm.precise_map(to_dev(vert_emb, torch.float64), to_dev(np.asarray(faces, dtype=np.int32)),
                                   to_dev(points_emb, torch.float64))
    face_match, bary = fm_d.cpu().numpy(), bary_d.cpu().numpy()
    P = barycentric_to_precise(faces, face_match, bary, n_vertices=vert_emb.shape[0])
    return (P, face_match, bary) if return_bary else P
