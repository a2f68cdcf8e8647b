from .convert import p2p_to_FM, mesh_p2p_to_FM, FM_to_p2p, mesh_FM_to_p2p  # noqa: F401
from .nn_utils import knn_query  # noqa: F401


from . import projection_utils  # noqa: F401


def mesh_FM_to_p2p_precise(FM_12, mesh1, mesh2, precompute_dmin=True, use_adj=True, batch_size=None, n_jobs=1,
                           verbose=False):
    """Barycentric "precise" map of mesh 2 onto mesh 1 (convert.py:186-231): (n2, n1) scipy csr matrix."""
    import numpy as np
    FM_12 = np.asarray(FM_12)
    k2, k1 = FM_12.shape
    if use_adj:
        emb1, emb2 = mesh1.eigenvectors[:, :k1], mesh2.eigenvectors[:, :k2] @ FM_12
    else:
        emb1, emb2 = mesh1.eigenvectors[:, :k1] @ FM_12.T, mesh2.eigenvectors[:, :k2]
    if mesh1.facelist is None:
        raise ValueError("the precise map needs the faces of mesh 1")
    return projection_utils.project_pc_to_triangles(emb1, mesh1.facelist, emb2, precompute_dmin=precompute_dmin,
                                                    batch_size=batch_size, n_jobs=n_jobs, verbose=verbose)
