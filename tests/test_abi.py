"""The C-ABI library loads and exports every symbol include/msfl.h declares (no GPU needed)."""
import ctypes as C
import os
import re

from msf_loam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "msfl.h")).read()
    return sorted(set(re.findall(r"\b(msfl_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert sorted(_lib.EXPORTS) == _declared()


def test_library_exports_every_symbol():
    lib = _lib.load_library()
    for name in _declared():
        assert hasattr(lib, name), name


def test_default_params_match_reference_constants():
    p = _lib.Params()
    _lib.load_library().msfl_default_params(C.byref(p))
    assert (p.min_range, p.scan_period, p.curvature_thresh, p.neighbor_gap_sq) == (0.3, 0.1, 0.1, 0.05)
    assert (p.n_sectors, p.n_sharp, p.n_less_sharp, p.n_flat) == (6, 2, 20, 4)
    assert (p.dist_sq_thresh, p.nearby_scan, p.min_correspondences) == (25.0, 2.5, 10)
    assert (p.knn_max_sq, p.line_eig_ratio, p.line_half_len, p.plane_tol) == (1.0, 3.0, 0.1, 0.2)
    assert (p.num_outer, p.max_num_iterations, p.huber_a) == (2, 6, 0.1)
    assert (p.initial_radius, p.min_relative_decrease) == (1e4, 1e-3)
    assert (p.function_tolerance, p.gradient_tolerance, p.parameter_tolerance) == (1e-6, 1e-10, 1e-8)


def test_struct_sizes_match_header():
    # sizes implied by include/msfl.h (natural alignment)
    assert C.sizeof(_lib.LmIter) == 48
    assert C.sizeof(_lib.LmLog) == 24 + 48 * _lib.MAX_ATTEMPTS
    assert C.sizeof(_lib.Stats) == 8 + 4 * 2 * _lib.MAX_OUTER + C.sizeof(_lib.LmLog) * _lib.MAX_OUTER
    assert C.sizeof(_lib.Cloud) == 48


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load_library()
    h = C.c_void_p()
    rc = lib.msfl_create(None, 0, C.byref(h))
    assert rc < 0 and b"no CPU fallback" in lib.msfl_last_error()
