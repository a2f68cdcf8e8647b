"""Drop-in mirror of the ``densematcher.pyFM`` surface that sits on the correspondence hot path
(SURVEY.md section 8b): ``spectral`` (knn_query, FM_to_p2p, p2p_to_FM, ...) and ``refine``
(zoomout_refine, icp_refine, ...).  numpy in / numpy out like the reference; the work happens in
libdm_b200.so on the current CUDA device."""
from . import spectral, refine, mesh, eval  # noqa: F401
from .mesh import TriMesh  # noqa: F401
from .functional import FunctionalMapping  # noqa: F401
