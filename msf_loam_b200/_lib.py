"""ctypes binding of libmsfl.so -- the C ABI declared in include/msfl.h.

There is no CPU fallback: if the shared library cannot be loaded (or built with nvcc), loading
raises; every compute call raises ``MsflError`` when the engine reports an error.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmsfl.so")

MSFL_OK = 0
MSFL_TOO_FEW = 1
MAX_OUTER = 4
MAX_ATTEMPTS = 16
N_STAGES = 4
NCCL_UNIQUE_ID_BYTES = 128
NO_FIELD = C.c_size_t(-1).value

# every symbol include/msfl.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "msfl_default_params", "msfl_last_error", "msfl_version", "msfl_abi_check", "msfl_create", "msfl_create_on_stream",
    "msfl_destroy", "msfl_sync", "msfl_stream", "msfl_launch_count", "msfl_set_profiling",
    "msfl_get_profile", "msfl_set_submap", "msfl_bcast_submap", "msfl_nccl_get_unique_id", "msfl_nccl_comm_init",
    "msfl_nccl_comm_destroy",
    "msfl_set_submap_device", "msfl_get_submap_device", "msfl_scan2map", "msfl_scan2map_batch",
    "msfl_scan2map_batch_device", "msfl_scan2map_batch_submit", "msfl_scan2map_batch_wait", "msfl_scan2map_deskew", "msfl_scan2map_deskew_batch", "msfl_associate_map", "msfl_scan2scan", "msfl_scan2scan_batch", "msfl_replay_batch", "msfl_associate_scan",
    "msfl_extract_features", "msfl_extract_features_batch", "msfl_register_and_match_batch", "msfl_voxel_grid", "msfl_accumulate", "msfl_map_create", "msfl_map_destroy",
    "msfl_cloud_from_pointcloud2", "msfl_map_insert", "msfl_map_surround", "msfl_map_size", "msfl_map_download", "msfl_set_submap_from_maps", "msfl_mapping_frame",
]


class MsflError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"msfl error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("min_range", C.c_double), ("scan_period", C.c_double), ("curvature_thresh", C.c_double),
        ("neighbor_gap_sq", C.c_double), ("n_sectors", C.c_int32), ("n_sharp", C.c_int32),
        ("n_less_sharp", C.c_int32), ("n_flat", C.c_int32),
        ("dist_sq_thresh", C.c_double), ("nearby_scan", C.c_double),
        ("min_correspondences", C.c_int32), ("_pad0", C.c_int32),
        ("knn_max_sq", C.c_double), ("line_eig_ratio", C.c_double), ("line_half_len", C.c_double),
        ("plane_tol", C.c_double),
        ("num_outer", C.c_int32), ("max_num_iterations", C.c_int32), ("huber_a", C.c_double),
        ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
        ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double), ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("max_consecutive_invalid_steps", C.c_int32), ("early_exit", C.c_int32),
        ("lm_cluster", C.c_int32), ("assoc_sorted", C.c_int32),
    ]


class Cloud(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("stride", C.c_size_t),
                ("off_xyz", C.c_size_t), ("off_intensity", C.c_size_t), ("off_ring", C.c_size_t)]


class LmIter(C.Structure):
    _fields_ = [("cost", C.c_double), ("cost_candidate", C.c_double), ("model_change", C.c_double),
                ("rho", C.c_double), ("radius", C.c_double), ("valid", C.c_int32), ("accepted", C.c_int32)]


class LmLog(C.Structure):
    _fields_ = [("n_attempts", C.c_int32), ("termination", C.c_int32), ("initial_cost", C.c_double),
                ("final_cost", C.c_double), ("it", LmIter * MAX_ATTEMPTS)]

    def as_dict(self):
        return {
            "n_attempts": self.n_attempts, "termination": self.termination,
            "initial_cost": self.initial_cost, "final_cost": self.final_cost,
            "iters": [{k: getattr(self.it[i], k) for k, _ in LmIter._fields_}
                      for i in range(self.n_attempts)],
        }


class Stats(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_outer", C.c_int32), ("n_edge", C.c_int32 * MAX_OUTER),
                ("n_plane", C.c_int32 * MAX_OUTER), ("lm", LmLog * MAX_OUTER)]

    def as_dict(self):
        return {
            "status": self.status, "n_outer": self.n_outer,
            "n_edge": list(self.n_edge)[: self.n_outer], "n_plane": list(self.n_plane)[: self.n_outer],
            "lm": [self.lm[i].as_dict() for i in range(self.n_outer)],
        }


class Deskew(C.Structure):
    _fields_ = [("sum_dt", C.POINTER(C.c_double)), ("delta_q", C.POINTER(C.c_double)),
                ("delta_p", C.POINTER(C.c_double)), ("n", C.c_int32), ("_pad", C.c_int32),
                ("velocity", C.c_double * 3), ("gravity", C.c_double * 3)]


class ChainCounts(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n_full", "n_sharp", "n_less_sharp", "n_flat", "n_less_flat",
                                          "n_corner_queries", "n_surf_queries")]


class Pc2Field(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint32), ("datatype", C.c_uint8), ("count", C.c_uint32)]


class Features(C.Structure):
    _fields_ = [("full_xyzi", C.POINTER(C.c_float)), ("full_ring", C.POINTER(C.c_uint16)),
                ("curvature", C.POINTER(C.c_float)), ("label", C.POINTER(C.c_int32)),
                ("idx_sharp", C.POINTER(C.c_int32)), ("idx_less_sharp", C.POINTER(C.c_int32)),
                ("idx_flat", C.POINTER(C.c_int32)), ("idx_less_flat", C.POINTER(C.c_int32)),
                ("n_full", C.c_int32), ("n_sharp", C.c_int32), ("n_less_sharp", C.c_int32),
                ("n_flat", C.c_int32), ("n_less_flat", C.c_int32)]


_lib = None


def load_library(build_if_needed: bool = True):
    """Loads libmsfl.so (building it in-tree with nvcc when missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    global LIB_PATH
    if os.environ.get("MSFL_LIB_PATH"):  # development: A/B an older build of the library against the tree's
        LIB_PATH, build_if_needed = os.environ["MSFL_LIB_PATH"], False
    if build_if_needed:
        from . import build as _build
        # a stale binary must never be loaded silently: the ctypes struct layouts below follow the CURRENT msfl.h
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m msf_loam_b200.build` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.msfl_last_error.restype = C.c_char_p
    lib.msfl_version.restype = C.c_char_p
    lib.msfl_stream.restype = C.c_void_p
    lib.msfl_stream.argtypes = [C.c_void_p]
    lib.msfl_launch_count.restype = C.c_uint64
    lib.msfl_launch_count.argtypes = [C.c_void_p]
    lib.msfl_destroy.restype = None
    lib.msfl_destroy.argtypes = [C.c_void_p]
    lib.msfl_default_params.restype = None
    lib.msfl_map_destroy.restype = None
    lib.msfl_map_destroy.argtypes = [C.c_void_p]
    # the .so reports the struct sizes it was compiled with; a mismatch means msfl.h and this binding disagree
    if os.environ.get("MSFL_LIB_PATH") and not hasattr(lib, "msfl_abi_check"):
        _lib = lib  # an older development build without the self-check
        return lib
    lib.msfl_abi_check.argtypes = [C.c_size_t] * 5
    rc = lib.msfl_abi_check(C.sizeof(Params), C.sizeof(Stats), C.sizeof(Cloud), C.sizeof(Features), C.sizeof(Deskew))
    if rc != MSFL_OK:
        raise ImportError(f"{LIB_PATH}: ABI mismatch with the ctypes binding ({lib.msfl_last_error().decode()}); "
                          "rebuild with `python -m msf_loam_b200.build --force`")
    _lib = lib
    return lib
