"""msfl_scan2map_batch_submit / _wait (up to three batches in flight: upload of batch k+1 overlaps the kernels of
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
    tm = e.scan2map_submit(pb, x0)  # MSFL_MAX_INFLIGHT = 3
    with pytest.raises(MsflError):
        e.scan2map_submit(pb, x0)
    out = np.zeros_like(x0)
    with pytest.raises(MsflError):
        e.scan2map_wait(tm + 5, out)
    e.scan2map_wait(t0, out)
    a = out.copy()
    e.scan2map_wait(t1, out)
    assert np.array_equal(a, out)
    e.scan2map_wait(tm, out)
    assert np.array_equal(a, out)
    with pytest.raises(MsflError):
        e.scan2map_wait(t1, out)  # already collected
    t2 = e.scan2map_submit(pb, x0)  # slots are free again
    e.scan2map_wait(t2, out)
    assert np.array_equal(a, out)
    e.close()


def test_xyz_only_and_pcl_layout_uploads_equal_the_float4_upload(vlp16_case):
    """The three host layouts of a batch -- packed float4 (DMA in place), packed xyz-only (12 B points, DMA in place +
    widening on the device), one 32-byte pcl::PointXYZI array per cloud (threaded host repack) -- give the same bits,
    through the asynchronous and the synchronous entry points."""
    from msf_loam_b200 import to_pcl
    e = Engine(default_params())
    e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
    (c, s, x0), = _batches(vlp16_case, 1, 48)
    cat_c, cat_s = np.concatenate(c), np.concatenate(s)
    co = np.concatenate([[0], np.cumsum([a.shape[0] for a in c])])
    so = np.concatenate([[0], np.cumsum([a.shape[0] for a in s])])
    c3, s3 = np.ascontiguousarray(cat_c[:, :3]), np.ascontiguousarray(cat_s[:, :3])
    layouts = {
        "float4": e.prepare_batch([cat_c[co[i]:co[i + 1]] for i in range(48)], [cat_s[so[i]:so[i + 1]] for i in range(48)]),
        "xyz": e.prepare_batch([c3[co[i]:co[i + 1]] for i in range(48)], [s3[so[i]:so[i + 1]] for i in range(48)]),
        "pcl": e.prepare_batch([to_pcl(a) for a in c], [to_pcl(a) for a in s]),
    }
    ref = None
    for name, pb in layouts.items():
        out = np.zeros_like(x0)
        e.scan2map_wait(e.scan2map_submit(pb, x0), out)
        x = x0.copy()
        e.scan2map_prepared(pb, x)
        assert np.array_equal(out, x), name
        if ref is None:
            ref = out
        assert np.array_equal(out, ref), name
    e.close()
