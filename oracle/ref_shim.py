"""Import the UNMODIFIED reference from /root/reference (authoring container only).

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_goldens.py`` to mint golden
vectors from the reference's own code and by optional CPU tests that are skipped
when ``/root/reference`` is absent (it does not exist on the GPU box).

Two native wheels the reference imports at module load are not installed
(``potpourri3d``, ``robust_laplacian``; densematcher/pyFM/mesh/trimesh.py:12-13,
mesh/geometry.py:6).  They are replaced by stub modules *in sys.modules only*;
the one stubbed function on the path, ``robust_laplacian.mesh_laplacian``
(trimesh.py:474), is answered with the reference's own vendored cotangent
Laplacian (mesh/laplacian.py:5-42, :88-140).  That changes only the eigenbasis,
which is a precomputed *input* of the hot path.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "densematcher", "pyFM"))


_loaded = {}


def load():
    """Returns a namespace with the reference's hot-path callables."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import scipy.sparse as sp

    sys.modules.setdefault("potpourri3d", types.ModuleType("potpourri3d"))
    rl = types.ModuleType("robust_laplacian")

    def mesh_laplacian(V, F, mollify_factor=1e-5):
        from densematcher.pyFM.mesh import laplacian as L
        return L.cotangent_weights(V, F), sp.csc_matrix(L.dia_area_mat(V, F))

    rl.mesh_laplacian = mesh_laplacian
    sys.modules.setdefault("robust_laplacian", rl)

    from densematcher.pyFM import spectral, refine
    from densematcher.pyFM.mesh import laplacian, TriMesh
    from densematcher.pyFM.functional import FunctionalMapping
    from densematcher.functional_map import compute_surface_map

    ns = types.SimpleNamespace(
        knn_query=spectral.knn_query, FM_to_p2p=spectral.FM_to_p2p, p2p_to_FM=spectral.p2p_to_FM,
        mesh_FM_to_p2p=spectral.mesh_FM_to_p2p, mesh_p2p_to_FM=spectral.mesh_p2p_to_FM,
        icp_refine=refine.icp_refine, zoomout_refine=refine.zoomout_refine,
        laplacian=laplacian, TriMesh=TriMesh, FunctionalMapping=FunctionalMapping,
        compute_surface_map=compute_surface_map, spectral=spectral, refine=refine,
    )
    _loaded["ns"] = ns
    return ns


class DuckMesh:
    """Stands in for a pytorch3d ``Meshes``: only ``verts_list``/``faces_list`` are read
    (densematcher/functional_map.py:17-18)."""

    def __init__(self, V, F):
        import torch
        self._v = torch.tensor(V, dtype=torch.float32)
        self._f = torch.tensor(F, dtype=torch.int64)

    def verts_list(self):
        return [self._v]

    def faces_list(self):
        return [self._f]
