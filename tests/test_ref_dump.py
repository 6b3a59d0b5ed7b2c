"""Pins the oracle to the REAL reference where a dump of it exists.

oracle/ref_harness/ builds `ref_dump` from the unmodified reference sources on a machine that has PCL + Ceres + Eigen
(not this image) and runs MappingScanMatcher::MatchScan2Map / OdometryScanMatcher::MatchScan2Scan on the arrays of
tests/golden/vlp16_golden.npz.  When its output has been committed as tests/golden/ref_dump_vlp16.txt this test compares
the oracle's poses and its Levenberg-Marquardt trace (cost, step quality, trust-region radius per iteration) with the
reference's own.  Until then the oracle is pinned to the reference's sources compiled against stand-in third-party headers
(oracle/_ref, tests/test_ref_*.py); what only this dump can pin is the real Ceres loop / FLANN / Eigen arithmetic."""
import os
import re

import numpy as np
import pytest

from msf_loam_b200 import synth as S

HERE = os.path.dirname(os.path.abspath(__file__))
DUMP = os.path.join(HERE, "golden", "ref_dump_vlp16.txt")
GOLDEN = os.path.join(HERE, "golden", "vlp16_golden.npz")


def parse_dump(text):
    """REF_MAP / REF_ODO pose lines + Ceres' progress tables (iter cost cost_change |gradient| |step| tr_ratio tr_radius ...)."""
    out = {"tables": []}
    table = None
    for line in text.splitlines():
        t = line.split()
        if t and t[0] in ("REF_MAP", "REF_ODO"):
            out[t[0]] = (int(t[1]), np.array(t[2:9], dtype=np.float64))
            continue
        if re.match(r"^\s*iter\s+cost\s+cost_change", line):
            table = []
            out["tables"].append(table)
            continue
        if table is not None and re.match(r"^\s*\d+\s+[-+0-9.eE]+", line) and len(t) >= 7:
            table.append([float(v) for v in t[:7]])  # iter cost cost_change |gradient| |step| tr_ratio tr_radius
        elif table is not None and not t:
            table = None
    return out


def test_parse_dump_format():
    sample = """iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time
   0  4.185660e+01    0.00e+00    1.09e+02   0.00e+00   0.00e+00  1.00e+04        0    5.34e-02    1.31e-01
   1  1.062590e+00    4.08e+01    5.36e+00   1.10e-01   9.95e-01  3.00e+04        1    5.29e-02    1.84e-01

REF_MAP 1 1.5 -2.0 0.25 0.0 0.0 0.1 0.99
REF_ODO 1 0.5 0.0 0.0 0.0 0.0 0.0 1.0
"""
    d = parse_dump(sample)
    assert d["REF_MAP"][0] == 1 and d["REF_MAP"][1].shape == (7,) and len(d["tables"]) == 1 and len(d["tables"][0]) == 2
    assert d["tables"][0][1][5] == pytest.approx(0.995) and d["tables"][0][1][6] == pytest.approx(3e4)


def test_oracle_against_reference_dump():
    if not os.path.exists(DUMP):
        pytest.skip("third-party numerics unpinned: no dump of the reference built against the real PCL + Ceres (build oracle/ref_harness where PCL + Ceres exist, "
                    "run it on tests/golden/vlp16_golden.npz and commit tests/golden/ref_dump_vlp16.txt)")
    g = np.load(GOLDEN)
    d = parse_dump(open(DUMP).read())
    ok, pose = d["REF_MAP"]
    dt, dr = S.pose_error(pose, g["pose_ref"])
    assert ok == 1 and dt <= 1e-6 and dr <= 1e-6, (dt, dr)  # north_star bound is 1e-4; same arithmetic should be far inside
    # one Ceres table per outer iteration: the cost column (row 0 = initial cost) and, for every later row, the step
    # quality and the radius AFTER that iteration; the oracle logs the radius used FOR each attempt
    trace = g["lm_trace_ref"]
    assert len(d["tables"]) >= trace.shape[0]
    for o in range(trace.shape[0]):
        tab = np.array(d["tables"][o])
        assert tab[0, 1] == pytest.approx(g["initial_cost_ref"][o], rel=1e-5)
        n = int(g["attempts_ref"][o])
        rows = tab[1:]
        assert rows.shape[0] in (n, n - 1)  # Ceres does not print the attempt that hits a tolerance
        for k in range(rows.shape[0]):
            assert rows[k, 5] == pytest.approx(trace[o, k, 2], rel=1e-2, abs=1e-3)  # tr_ratio is printed with 3 digits
            if k + 1 < n:
                assert rows[k, 6] == pytest.approx(trace[o, k + 1, 3], rel=1e-2)
    ok, pose = d["REF_ODO"]
    dt, dr = S.pose_error(pose, g["odo_pose"])
    assert ok == (1 if int(g["odo_rc"]) == 0 else 0) and dt <= 1e-6 and dr <= 1e-6, (dt, dr)
