"""Generates tests/golden/vlp16_golden.npz from the CPU oracle (the reference itself cannot be built
or run here -- no PCL/Ceres/Eigen/ROS -- so these are ORACLE outputs, "parity unpinned").

    python tests/golden/make_golden.py

Contents: one raw VLP-16 scan with its expected feature index lists, a 5-scan submap, one query scan,
the initial guess, and the oracle's association / LM trace / final pose for both schedules.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
from conftest import make_map_case  # noqa: E402


def main():
    case = make_map_case()
    q = case["queries"][0]
    xyzi, ring = q["raw"]
    P = O.default_params()
    f = O.extract_features(P, xyzi, ring, None)
    corr, ne, npl, kidx = O.associate_map(P, case["map_corner"], case["map_surf"], q["corner"], q["surf"], q["init"])
    out = {
        "raw_xyzi": xyzi, "raw_ring": ring,
        "feat_idx_sharp": f["idx_sharp"], "feat_idx_less_sharp": f["idx_less_sharp"],
        "feat_idx_flat": f["idx_flat"], "feat_idx_less_flat": f["idx_less_flat"],
        "feat_curvature": f["curvature"], "feat_time": f["full"][:, 3],
        "map_corner": case["map_corner"], "map_surf": case["map_surf"],
        "scan_corner": q["corner"], "scan_surf": q["surf"], "init": q["init"], "gt": q["gt"],
        "knn_idx": kidx, "n_edge0": ne, "n_plane0": npl, "corr0": corr,
    }
    # scan-to-scan pair (BASELINE config 1): features of two consecutive query scans, identity initial guess
    f0, f1 = case["queries"][0]["features"], case["queries"][1]["features"]
    odo = {"odo_last_corner": f0["full"][f0["idx_less_sharp"]], "odo_last_corner_ring": f0["ring"][f0["idx_less_sharp"]],
           "odo_last_surf": f0["full"][f0["idx_less_flat"]], "odo_last_surf_ring": f0["ring"][f0["idx_less_flat"]],
           "odo_curr_sharp": f1["full"][f1["idx_sharp"]], "odo_curr_flat": f1["full"][f1["idx_flat"]],
           "odo_init": np.array([0, 0, 0, 0, 0, 0, 1.0])}
    rc, x, logs, counts, assoc = O.scan2scan(P, odo["odo_last_corner"], odo["odo_last_corner_ring"], odo["odo_last_surf"],
                                            odo["odo_last_surf_ring"], odo["odo_curr_sharp"], odo["odo_curr_flat"], odo["odo_init"])
    out.update(odo)
    out.update(odo_rc=rc, odo_pose=x, odo_counts=counts, odo_assoc=assoc,
               odo_final_cost=np.array([l["final_cost"] for l in logs]))
    out["lm_trace_ref"] = None  # filled below
    for name, over in (("ref", {}), ("fixed10", {"early_exit": 0, "max_num_iterations": 5})):
        Pn = O.default_params(**over)
        x, logs, counts = O.scan2map(Pn, case["map_corner"], case["map_surf"], q["corner"], q["surf"], q["init"])
        out[f"pose_{name}"] = x
        out[f"counts_{name}"] = counts
        out[f"attempts_{name}"] = np.array([l["n_attempts"] for l in logs])
        out[f"final_cost_{name}"] = np.array([l["final_cost"] for l in logs])
        out[f"accepted_{name}"] = np.array([[it["accepted"] for it in l["iters"]] + [-1] * (16 - l["n_attempts"]) for l in logs])
        # Ceres-style trace per attempt: cost before, candidate cost, rho (tr_ratio), radius (tr_radius)
        out[f"lm_trace_{name}"] = np.array([[[it["cost"], it["cost_candidate"], it["rho"], it["radius"]] for it in l["iters"]] +
                                            [[np.nan] * 4] * (16 - l["n_attempts"]) for l in logs])
        out[f"initial_cost_{name}"] = np.array([l["initial_cost"] for l in logs])
    path = os.path.join(ROOT, "tests", "golden", "vlp16_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
