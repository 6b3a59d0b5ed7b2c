"""msf_loam_b200 -- Blackwell-native (sm_100a) LOAM scan-matching engine.

One hot path of kekeliu-whu/MSF_LOAM (feature extraction -> scan-to-scan -> scan-to-map), written
as CUDA kernels behind the C ABI in ``include/msfl.h`` (``libmsfl.so``, built in-tree by
``msf_loam_b200.build``).  ``engine`` mirrors the reference's matcher interface on top of it;
``synth`` generates seeded synthetic LiDAR scenes.  There is no CPU fallback.
"""
from .engine import (Engine, HybridGrid, set_submap_from_maps, mapping_frame, MappingScanMatcher, OdometryScanMatcher, ScanRegistration,
                     TimestampedPointCloud, default_params, to_pcl)
from ._lib import MsflError, Params, load_library

__all__ = ["Engine", "HybridGrid", "set_submap_from_maps", "mapping_frame", "MappingScanMatcher", "OdometryScanMatcher", "ScanRegistration",
           "TimestampedPointCloud", "default_params", "to_pcl", "MsflError", "Params", "load_library"]
