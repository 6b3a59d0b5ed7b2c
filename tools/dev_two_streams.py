"""Developer experiment (NOT bench.py): does running the association of one half-batch concurrently with the LM solve of
the other (two engines = two streams with their own scratch) beat one engine on the whole batch?  Inputs from the bench
workload (VLP-16, 2048 scans)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from msf_loam_b200 import Engine, default_params

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
over = dict(bench.OVER)
traj, scans = bench.raw_scans("vlp16", 32, 8)
_e = Engine(default_params(**over))
map_corner, map_surf, queries, _ = bench.build_case(lambda x, r: _e.extract_features(x, r, None), _e.voxel_grid, "vlp16", traj, scans)
_e.close()
qc, c_off, qs, s_off, inits = bench.assemble_batch(queries[:32], B, seed=1000)


def half(lo, hi):
    c0, c1, s0, s1 = int(c_off[lo]), int(c_off[hi]), int(s_off[lo]), int(s_off[hi])
    return {"B": hi - lo, "d_c": torch.from_numpy(qc[c0:c1].copy()).to(dev), "d_s": torch.from_numpy(qs[s0:s1].copy()).to(dev),
            "d_co": torch.from_numpy((c_off[lo:hi + 1] - c0).astype(np.int32)).to(dev),
            "d_so": torch.from_numpy((s_off[lo:hi + 1] - s0).astype(np.int32)).to(dev), "nc": c1 - c0, "ns": s1 - s0,
            "p0": torch.from_numpy(inits[lo:hi].copy()).to(dev)}


def make(parts):
    out = []
    for lo, hi in parts:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            e = Engine(default_params(**over), stream=st.cuda_stream)
            e.set_submap(map_corner, map_surf)
            h = half(lo, hi)
            h["p"] = h["p0"].clone()
        out.append((st, e, h))
    torch.cuda.synchronize()
    return out


def run(ctx, sleep_cycles=0):
    def one():
        for st, e, h in ctx:
            with torch.cuda.stream(st):
                h["p"].copy_(h["p0"])
                e.scan2map_batch_device(h["B"], h["d_c"].data_ptr(), h["d_co"].data_ptr(), h["nc"], h["d_s"].data_ptr(),
                                        h["d_so"].data_ptr(), h["ns"], h["p"].data_ptr())
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    main = ctx[0][0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for st, _, _ in ctx[1:]:
        st.wait_stream(main)
    ev0.record(main)
    for st, _, _ in ctx[1:]:
        st.wait_event(ev0)
        if sleep_cycles:
            with torch.cuda.stream(st):
                torch.cuda._sleep(sleep_cycles)
    for _ in range(steps):
        one()
    for st, _, _ in ctx[1:]:
        main.wait_stream(st)
    ev1.record(main)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


one = make([(0, B)])
ms1 = run(one)
print(f"one engine, {B} scans: {ms1:.3f} ms/step = {B / ms1 * 1e3:.0f} scans/s")
ref = one[0][2]["p"].cpu().numpy().copy()
for st, e, h in one:
    e.close()
two = make([(0, B // 2), (B // 2, B)])
for cyc in (0, 1_000_000, 1_500_000, 2_000_000):
    ms2 = run(two, cyc)
    print(f"two engines x {B // 2} scans, second delayed by {cyc / 1.9e6:.2f} ms: {ms2:.3f} ms/step = {B / ms2 * 1e3:.0f} scans/s")
got = np.concatenate([two[0][2]["p"].cpu().numpy(), two[1][2]["p"].cpu().numpy()])
print("poses bit-identical to the single-engine run:", np.array_equal(got, ref))
four = make([(i * B // 4, (i + 1) * B // 4) for i in range(4)])
ms4 = run(four, 700_000)
print(f"four engines x {B // 4}: {ms4:.3f} ms/step = {B / ms4 * 1e3:.0f} scans/s")
