"""msfl_scan2map_batch_submit / _wait (two batches in flight: upload of batch k+1 overlaps the kernels of
batch k) must give bit-identical poses to the synchronous msfl_scan2map_batch and stay within tolerance of the oracle."""
import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import Engine, MsflError, default_params
from msf_loam_b200 import synth as S

pytestmark = pytest.mark.gpu


def _batches(case, n_batches, B):
    rng = np.random.default_rng(7)
    out = []
    for _ in range(n_batches):
        qs = [case["queries"][i % len(case["queries"])] for i in range(B)]
        inits = np.stack([S.perturb_pose(q["gt"], rng) for q in qs])
        out.append(([q["corner"] for q in qs], [q["surf"] for q in qs], inits))
    return out


def test_pipelined_batches_equal_synchronous_bitwise(vlp16_case):
    e = Engine(default_params())
    e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
    batches = _batches(vlp16_case, 5, 24)
    sync = []
    for c, s, x0 in batches:
        x = x0.copy()
        e.scan2map_prepared(e.prepare_batch(c, s), x)
        sync.append(x)
    prepared = [e.prepare_batch(c, s) for c, s, _ in batches]
    outs = [np.zeros_like(b[2]) for b in batches]
    tickets = [e.scan2map_submit(prepared[0], batches[0][2])]
    for k in range(1, len(batches)):
        tickets.append(e.scan2map_submit(prepared[k], batches[k][2]))  # two in flight
        e.scan2map_wait(tickets[k - 1], outs[k - 1])
    e.scan2map_wait(tickets[-1], outs[-1])
    for a, b in zip(sync, outs):
        assert np.array_equal(a, b)
    # and the oracle agrees on one scan of the last batch
    P = O.default_params()
    c, s, x0 = batches[-1]
    ref, _, _ = O.scan2map(P, vlp16_case["map_corner"], vlp16_case["map_surf"], c[3], s[3], x0[3])
    dt, dr = S.pose_error(outs[-1][3], ref)
    assert dt <= 1e-4 and dr <= 1e-4
    e.close()


def test_third_submit_without_wait_is_refused_and_bad_ticket(vlp16_case):
    e = Engine(default_params())
    e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
    (c, s, x0), = _batches(vlp16_case, 1, 4)
    pb = e.prepare_batch(c, s)
    t0 = e.scan2map_submit(pb, x0)
    t1 = e.scan2map_submit(pb, x0)
    with pytest.raises(MsflError):
        e.scan2map_submit(pb, x0)
    out = np.zeros_like(x0)
    with pytest.raises(MsflError):
        e.scan2map_wait(t1 + 5, out)
    e.scan2map_wait(t0, out)
    a = out.copy()
    e.scan2map_wait(t1, out)
    assert np.array_equal(a, out)
    with pytest.raises(MsflError):
        e.scan2map_wait(t1, out)  # already collected
    t2 = e.scan2map_submit(pb, x0)  # slots are free again
    e.scan2map_wait(t2, out)
    assert np.array_equal(a, out)
    e.close()
