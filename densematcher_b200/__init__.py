"""densematcher_b200 -- B200-native (sm_100a) implementation of DenseMatcher's correspondence hot path.

    densematcher_b200.nn            fused similarity + argmax (torch tensors)
    densematcher_b200.fm            functional-map stages (projection, solve, FM<->p2p, ZoomOut, ICP)
    densematcher_b200.pipeline      batched pipeline, host-buffer entry, sharding across ranks
    densematcher_b200.pyFM          drop-in mirror of densematcher.pyFM.{spectral,refine} (numpy in/out)
    densematcher_b200.functional_map  drop-in compute_surface_map

All compute happens in libdm_b200.so (include/dm_b200.h); importing the package does not require a GPU,
calling it does -- there is no CPU fallback.
"""
__version__ = "0.1.0"
