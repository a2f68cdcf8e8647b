from .convert import p2p_to_FM, mesh_p2p_to_FM, FM_to_p2p, mesh_FM_to_p2p  # noqa: F401
from .nn_utils import knn_query  # noqa: F401
from .shape_difference import area_SD, conformal_SD, compute_SD  # noqa: F401


def mesh_FM_to_p2p_precise(*args, **kwargs):
    """Barycentric "precise" map (convert.py:185-229, projection_utils.py): only reached with ``compute_extra=True``;
    not implemented."""
    raise NotImplementedError("mesh_FM_to_p2p_precise (barycentric precise map) is not implemented")
