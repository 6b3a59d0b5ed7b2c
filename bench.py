#!/usr/bin/env python
"""bench.py -- scans/sec of the LOAM scan-to-map hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

One "step" = one pass of MatchScan2Map over a batch of independent synthetic scans against one shared submap.
The headline line is BASELINE config 2 (VLP-16 scan vs 5-scan corner/surf submap, 2 outer iterations x 5 LM
attempts); at N>1 the scans are sharded over the GPUs and every step the owner rank re-indexes the submap and
broadcasts it WITH its cell index through the C ABI (msfl_bcast_submap, NCCL), the other ranks adopt it (config 4).
The same JSON line carries a `workloads` record with BASELINE configs 3 (HDL-64E), 4 (one scan per GPU, broadcast
+ adoption + solve latency) and 5 (OS1-128 vs a ~1 M-point 50-scan submap, batch 1 / 8 / 64 per GPU).
Inputs are produced by the product path itself (ray-cast -> CUDA feature extraction -> CUDA VoxelGrid); the CPU
oracle is only executed for `cpu_baseline`, the pose checks, and `--impl reference`.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from msf_loam_b200 import synth as S  # noqa: E402

# name: sensor, scene, submap scans, trajectory (step m, yaw deg, jitter deg, start), description
WORKLOADS = {
    "vlp16": dict(sensor="vlp16", scene="room40", map_scans=5, traj=dict(step=0.5, yaw_deg=1.0), seed0=100,
                  desc="VLP-16 scan-to-map: 29k-pt scan vs 5-scan corner/surf submap, 10 LM iters"),
    "hdl64": dict(sensor="hdl64", scene="room80", map_scans=5, traj=dict(step=0.5, yaw_deg=1.0), seed0=200,
                  desc="HDL-64E KITTI-shape scan (~130k pts, 64 rings) scan-to-map vs 5-scan submap, 10 LM iters"),
    "os1-128": dict(sensor="os1-128", scene="hall300", map_scans=50, seed0=400,
                    traj=dict(step=5.0, yaw_deg=0.0, jitter_deg=0.05, start=(-140.0, -3.0, 1.5)),
                    desc="OS1-128-shape 260k-pt scan vs ~1M-pt (50-scan) submap, 10 LM iters"),
}
SIGMA = 0.01
K_OUTER, L_ATTEMPTS = 2, 5  # "10 LM iters" = 2 outer x 5 attempts, fixed count
OVER = {"early_exit": 0, "max_num_iterations": L_ATTEMPTS, "num_outer": K_OUTER}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vlp16", choices=sorted(WORKLOADS), help="headline workload")
    ap.add_argument("--batch", type=int, default=2048, help="scans per GPU per step (headline workload)")
    ap.add_argument("--distinct", type=int, default=32, help="distinct query scans (replicated to fill the batch)")
    ap.add_argument("--cpu-sample", type=int, default=1536, help="scans timed on the CPU oracle for cpu_baseline")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / pose-check legs (development)")
    ap.add_argument("--no-workloads", action="store_true", help="headline workload only (development, profiling)")
    ap.add_argument("--only-device", action="store_true", help="device-resident leg only (profiling under ncu)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# workload generation
# ------------------------------------------------------------------------------------------------
def _raycast_job(a):
    workload, k, pose = a
    w = WORKLOADS[workload]
    return S.raycast_scan(S.make_scene(w["scene"]), w["sensor"], pose, seed=w["seed0"] + k, sigma=SIGMA)


def raw_scans(workload, n_distinct, n_workers=1):
    """Trajectory + raw ray-cast scans (numpy; done before any CUDA context exists so a fork pool is safe)."""
    w = WORKLOADS[workload]
    traj = S.trajectory(w["map_scans"] + n_distinct, **w["traj"])
    jobs = [(workload, k, traj[k]) for k in range(len(traj))]
    if n_workers > 1 and len(jobs) > 8:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(n_workers, len(jobs))) as pool:
            scans = pool.map(_raycast_job, jobs, chunksize=1)
    else:
        scans = [_raycast_job(j) for j in jobs]
    return traj, scans


def build_case(extract, voxel, workload, traj, scans):
    """submap + query features; `extract` / `voxel` are the engine's (product path) or the oracle's (reference arm)."""
    n_map = WORKLOADS[workload]["map_scans"]
    mc, ms, queries, n_pts = [], [], [], []
    for k, (xyzi, ring) in enumerate(scans):
        f = extract(xyzi, ring)
        n_pts.append(f["full"].shape[0])
        corner, surf = f["full"][f["idx_less_sharp"]], f["full"][f["idx_less_flat"]]
        if k < n_map:
            mc.append(S.transform_cloud(traj[k], corner))
            ms.append(S.transform_cloud(traj[k], surf))
        else:
            queries.append((voxel(corner, 0.2), voxel(surf, 0.4), traj[k]))
    return voxel(np.concatenate(mc), 0.2), voxel(np.concatenate(ms), 0.4), queries, int(np.mean(n_pts))


def assemble_batch(queries, B, seed):
    """B scans: distinct scan (i mod D) with its own seeded initial-guess perturbation (0.10 m, 1 deg)."""
    rng = np.random.default_rng(seed)
    D = len(queries)
    corners = [queries[i % D][0] for i in range(B)]
    surfs = [queries[i % D][1] for i in range(B)]
    inits = np.stack([S.perturb_pose(queries[i % D][2], rng) for i in range(B)])
    c_off = np.concatenate([[0], np.cumsum([c.shape[0] for c in corners])]).astype(np.int32)
    s_off = np.concatenate([[0], np.cumsum([c.shape[0] for c in surfs])]).astype(np.int32)
    return np.concatenate(corners), c_off, np.concatenate(surfs), s_off, inits


def algorithmic_bytes(N, M, K=K_OUTER, L=L_ATTEMPTS):
    """bytes = K [16 M + 48 N + 48 (1+L) N]; also the per-launch split."""
    assoc = 16 * M + 48 * N            # queries + submap once + correspondences written
    solve = 48 * (1 + L) * N           # one fused residual/Jacobian/cost sweep per evaluation point
    return K * (assoc + solve), assoc, solve


def workload_config(workload, B, queries, n_full, n_q, Mc, Ms):
    """The workload description shared VERBATIM by the CUDA arm and the reference arm (the driver compares it)."""
    return {"workload": WORKLOADS[workload]["desc"], "sensor": workload, "scans_per_gpu_per_step": B,
            "distinct_scans": len(queries), "points_per_scan": n_full, "queries_per_scan": round(n_q / B, 1),
            "submap_points": {"corner": Mc, "surf": Ms}, "submap_scans": WORKLOADS[workload]["map_scans"],
            "outer_iterations": K_OUTER, "lm_attempts_per_outer": L_ATTEMPTS, "early_exit": False,
            "range_noise_sigma_m": SIGMA, "initial_guess_error": "0.10 m, 1.0 deg (seeded)"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        if os.environ.get("MSFL_BENCH_NO_SAMPLER"):  # development: is the nvidia-smi poll perturbing the host-side timing?
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v == "Active":
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # "under load" = samples in the upper half of the observed power range
            pw = np.array(power)
            load = pw >= (pw.min() + 0.5 * (pw.max() - pw.min())) if pw.max() > pw.min() else np.ones_like(pw, bool)
            out.update(sm_mhz=float(np.median(np.array(sm)[load])), sm_max_mhz=float(max(smax)),
                       reasons=sorted(reasons), samples=len(sm), samples_under_load=int(load.sum()),
                       power_w_max=float(pw.max()))
        return out


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, batch):
    """dram read+write bytes per launch of the stage kernels from the committed ncu capture of this command
    (profiles/roofline_traffic.json; ncu cannot run inside the timed process)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(f"{workload}:{batch}")
    except Exception:
        return None


def numa_bind(local_rank):
    """Binds this rank to the NUMA node of its GPU when the box exposes more than one node (pinned buffers are then
    allocated node-local by first touch).  Returns a description for the JSON record."""
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        nodes = []
    info = {"numa_nodes_visible": len(nodes), "cpus": len(os.sched_getaffinity(0)), "bound": None}
    if len(nodes) < 2:
        return info
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id  # type: ignore[attr-defined]
    except Exception:
        bdf = None
    try:
        q = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                           capture_output=True, text=True).stdout.strip().lower()
        if q.startswith("00000000:"):
            q = "0000:" + q[9:]
        with open(f"/sys/bus/pci/devices/{q}/numa_node") as f:
            node = int(f.read().strip())
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"] = {"node": node, "cpus": len(cpus)}
    except Exception as ex:  # topology not readable inside the container: stay unbound
        info["error"] = str(ex)[:80]
    return info


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
STAGE_NAMES = {0: "k_knn5_fit (5-NN search over the submap cell index + line / plane fit)", 1: "k_lm_solve (residual/Jacobian/6x6 LM)",
               2: "k_transform_keys + counting sort of the cell keys", 3: "k_fit_qr_list (Householder fallback of declined plane fits)"}


class Ctx:
    pass


def run_workload(cx, workload, B, n_distinct, steps, warmup, full, lm_cluster=0, sampler=True, e2e=True, cpu_scans=0,
                 bcast=True, remap=False):
    """One workload on every rank: device-resident leg (`value`), host-buffer legs (`e2e`), roofline, pose check."""
    import torch
    import torch.distributed as dist
    from msf_loam_b200 import Engine, default_params, to_pcl

    args, rank, world, dev = cx.args, cx.rank, cx.world, cx.dev
    over = dict(OVER)
    if lm_cluster:
        over["lm_cluster"] = lm_cluster
    stream = torch.cuda.Stream(device=dev)
    rec = {}
    with torch.cuda.stream(stream):
        eng = Engine(default_params(**over), device=cx.local_rank, stream=stream.cuda_stream)
        traj, scans = cx.raw[workload]
        map_corner, map_surf, queries, n_full = build_case(lambda x, r: eng.extract_features(x, r, None), eng.voxel_grid,
                                                           workload, traj, scans)
        queries = queries[:n_distinct]
        if B < len(queries):  # fewer scans than distinct ones (config 4): every rank takes its own
            r = (rank * B) % len(queries)
            queries = queries[r:] + queries[:r]
        qc, c_off, qs, s_off, inits = assemble_batch(queries, B, seed=1000 + rank)
        Mc, Ms = map_corner.shape[0], map_surf.shape[0]
        n_q = int(c_off[-1] + s_off[-1])
        # the submap is owned by rank 0; every step rank 0 indexes a new map version and broadcasts it with its index
        # (C ABI, NCCL), the other ranks adopt it without building anything
        t_mc, t_ms = torch.from_numpy(map_corner).to(dev), torch.from_numpy(map_surf).to(dev)
        comm = None
        if world > 1 and bcast:
            from msf_loam_b200 import sharding
            comm = sharding.make_submap_comm(eng, device=dev)
        if rank == 0 or comm is None:
            eng.set_submap_device(t_mc.data_ptr(), Mc, t_ms.data_ptr(), Ms)
        if comm is not None:
            eng.bcast_submap(comm, 0)
        d_qc, d_qs = torch.from_numpy(qc).to(dev), torch.from_numpy(qs).to(dev)
        d_co, d_so = torch.from_numpy(c_off).to(dev), torch.from_numpy(s_off).to(dev)
        d_p0 = torch.from_numpy(inits).to(dev)
        d_p = d_p0.clone()

        def new_map_version():
            if comm is None:
                if remap:  # single GPU: the new map version is indexed in place
                    eng.set_submap_device(t_mc.data_ptr(), Mc, t_ms.data_ptr(), Ms)
                return
            if rank == 0:
                eng.set_submap_device(t_mc.data_ptr(), Mc, t_ms.data_ptr(), Ms)
            eng.bcast_submap(comm, 0)

        def step_device():
            new_map_version()
            d_p.copy_(d_p0)
            eng.scan2map_batch_device(B, d_qc.data_ptr(), d_co.data_ptr(), int(c_off[-1]), d_qs.data_ptr(),
                                      d_so.data_ptr(), int(s_off[-1]), d_p.data_ptr())

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        smp = ClockSampler(cx.local_rank) if sampler else None
        for _ in range(max(warmup, 3)):
            step_device()
        barrier()
        eng.get_profile()
        eng.set_profiling(True)
        launches0 = eng.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            step_device()
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        launches = eng.launch_count - launches0
        stage_ms, stage_cnt = eng.get_profile()
        eng.set_profiling(False)
        poses_dev = d_p.cpu().numpy()

        e2e_ms = e2e_pcl_ms = e2e_sync_ms = e2e_f4_ms = None
        if e2e and not args.only_device:
            # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ---------------------
            h_qc, h_qs = torch.from_numpy(qc).pin_memory(), torch.from_numpy(qs).pin_memory()
            hc, hs = h_qc.numpy(), h_qs.numpy()
            packed = eng.prepare_batch([hc[c_off[i]:c_off[i + 1]] for i in range(B)],
                                       [hs[s_off[i]:s_off[i + 1]] for i in range(B)])
            # xyz-only clouds (12 B points) in pinned memory: all the LiDAR-only matcher reads of a query
            h_qc3 = torch.from_numpy(np.ascontiguousarray(qc[:, :3])).pin_memory()
            h_qs3 = torch.from_numpy(np.ascontiguousarray(qs[:, :3])).pin_memory()
            hc3, hs3 = h_qc3.numpy(), h_qs3.numpy()
            packed3 = eng.prepare_batch([hc3[c_off[i]:c_off[i + 1]] for i in range(B)],
                                        [hs3[s_off[i]:s_off[i + 1]] for i in range(B)])
            # the layout the reference's adapter passes: one pcl::PointCloud<pcl::PointXYZI> per cloud (32 B points,
            # pageable memory, separate allocations); the library repacks them into its pinned slot on host threads
            D = len(queries)  # every scan of the batch gets its own arrays, as B independent clouds would have
            pcl = eng.prepare_batch([to_pcl(queries[i % D][0]) for i in range(B)], [to_pcl(queries[i % D][1]) for i in range(B)])
            h_p = inits.copy()
            eng.scan2map_prepared(packed, h_p)  # synchronous call (the ROS drop-in form)
            assert np.array_equal(h_p, poses_dev), "host-buffer and device-resident paths disagree"
            h_p[:] = inits
            eng.scan2map_prepared(pcl, h_p)
            assert np.array_equal(h_p, poses_dev), "PCL-layout host path and device-resident path disagree"
            h_out = [np.zeros_like(inits), np.zeros_like(inits)]

            def pipelined(prepared):
                """K steps through msfl_scan2map_batch_submit / _wait, three in flight: the upload of step k+1 overlaps
                the kernels of step k; every step's inputs come from host memory, every step's poses return to it.
                The map is frozen during the replay (broadcast and adopted once, before the loop): a new map version
                costs a host synchronisation on the adopting ranks, which would serialise the batches in flight --
                its cost is in the device-timed `value` and in workloads.config4."""
                depth = 3  # MSFL_MAX_INFLIGHT: repack of k+2 | upload of k+1 | kernels of k
                tickets = []
                for i in range(steps):
                    if len(tickets) == depth:
                        eng.scan2map_wait(tickets.pop(0), h_out[i & 1])
                    tickets.append(eng.scan2map_submit(prepared, inits))
                for j, tk in enumerate(tickets):
                    eng.scan2map_wait(tk, h_out[j & 1])
                torch.cuda.synchronize(dev)

            def timed(fn, reps=3):
                runs = []
                for _ in range(reps):  # the median run is reported (host-side timing of a short region is noisy)
                    barrier()
                    t0 = time.perf_counter()
                    fn()
                    runs.append(time.perf_counter() - t0)
                return sorted(runs)[len(runs) // 2] * 1e3

            pipelined(packed3)
            e2e_ms = timed(lambda: pipelined(packed3))
            assert np.array_equal(h_out[0], poses_dev) and (steps < 2 or np.array_equal(h_out[1], poses_dev)), \
                "pipelined xyz-only host-buffer path and device-resident path disagree"
            pipelined(packed)
            e2e_f4_ms = timed(lambda: pipelined(packed))
            assert np.array_equal(h_out[0], poses_dev), "pipelined float4 host-buffer path and device-resident path disagree"
            pipelined(pcl)
            e2e_pcl_ms = timed(lambda: pipelined(pcl))
            assert np.array_equal(h_out[0], poses_dev), "pipelined PCL-layout path and device-resident path disagree"

            def sync_calls():  # the synchronous single-call form, for the record
                for _ in range(steps):
                    h_p[:] = inits
                    eng.scan2map_prepared(packed, h_p)
            e2e_sync_ms = timed(sync_calls, reps=1)
        # the clock sampler spans the device-timed steps AND the host-buffer legs (the same kernels under the same load)
        rec["clocks"] = smp.stop() if smp else None
        if comm is not None:
            eng.sync()
            eng.nccl_comm_destroy(comm)
        eng.close()

    # max over ranks
    vals = [ms_total, e2e_ms or 0.0, e2e_pcl_ms or 0.0, e2e_sync_ms or 0.0, e2e_f4_ms or 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, e2e_pcl_ms, e2e_sync_ms, e2e_f4_ms = (float(v) for v in t)
    ms_per_step = ms_total / steps
    rate = lambda ms: round(world * B * steps / (ms * 1e-3), 1) if ms else None  # noqa: E731
    rec.update(value=rate(ms_total), ms_per_step=round(ms_per_step, 4), gpu_launches=int(launches))
    if rank != 0:
        return rec
    peak, peak_src = measured_peak()
    total_bytes, assoc_bytes, solve_bytes = algorithmic_bytes(n_q, Mc + Ms)
    per_launch_ms = {s: stage_ms[s] / stage_cnt[s] for s in range(len(stage_ms)) if stage_cnt[s]}
    dom = max(per_launch_ms, key=lambda s: stage_ms[s])  # dominant kernel = the stage with the largest share of the step
    dom_bytes = assoc_bytes if dom in (0, 2, 3) else solve_bytes
    achieved = dom_bytes / (per_launch_ms[dom] * 1e-3) / 1e9
    tr = ncu_traffic(workload, B) or {}
    rec["roofline"] = {
        "bound": "hbm", "kernel": STAGE_NAMES[dom], "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 5), "traffic": tr.get(str(dom)), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": round(per_launch_ms[dom], 4),
        "stage_ms_per_step": {STAGE_NAMES[s]: round(stage_ms[s] / steps, 4) for s in per_launch_ms},
        "stage_share": {STAGE_NAMES[s]: round(stage_ms[s] / sum(stage_ms), 3) for s in per_launch_ms},
        "whole_step_GBps": round(total_bytes / (ms_per_step * 1e-3) / 1e9, 2),
        "whole_step_frac": round(total_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4),
        "traffic_source": tr.get("source"),
        "note": "achieved = SURVEY 8d algorithmic bytes of that pass / CUDA-event time of the stage; DESIGN.md section 4"}
    rec["config"] = workload_config(workload, B, queries, n_full, n_q, Mc, Ms)
    rec["impl_notes"] = {
        "parallelism": f"scan-sharded x{world}" + (", per step: rank 0 re-indexes the submap, msfl_bcast_submap "
                                                    "(NCCL, points + cell index), other ranks adopt" if world > 1 and bcast else ""),
        "lm_cluster": lm_cluster or 1,
        "l2": "per-step inputs + correspondences (%.0f MB) vs the 126 MB L2" % ((n_q * 16 + n_q * 48) / 1e6)}
    if e2e_ms:
        h2d = int(n_q * 12 + (2 * (B + 1)) * 4 + B * 56)
        h2d_f4 = int(n_q * 16 + (2 * (B + 1)) * 4 + B * 56)
        rec["e2e"] = {"value": rate(e2e_ms), "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(B * 56),
                      "api": "msfl_scan2map_batch_submit/_wait, 3 batches in flight, xyz-only clouds (12 B points: all the "
                             "LiDAR-only matcher reads of a query) in pinned host memory, DMA'd in place and widened on the "
                             "device; median of 3 runs of K steps",
                      "h2d_GBps_per_rank": round(h2d * steps / (e2e_ms * 1e-3) / 1e9, 2),
                      "float4_value": rate(e2e_f4_ms), "float4_h2d_bytes_per_step": h2d_f4,
                      "float4_h2d_GBps_per_rank": round(h2d_f4 * steps / (e2e_f4_ms * 1e-3) / 1e9, 2) if e2e_f4_ms else None,
                      "pcl_layout_value": rate(e2e_pcl_ms),
                      "pcl_layout": "same calls fed one pcl::PointXYZI-layout array per cloud (32 B points, pageable, "
                                    "separate allocations): repacked into the pinned slot by %s host threads" %
                                    os.environ.get("MSFL_PACK_THREADS", "up to 16"),
                      "synchronous_call_value": rate(e2e_sync_ms)}
    if cpu_scans and not args.no_cpu:
        rec["cpu_baseline"], rec["pose_err_vs_oracle"] = cpu_baseline(map_corner, map_surf, qc, c_off, qs, s_off, inits,
                                                                        poses_dev, min(cpu_scans, B), over)
    return rec


def single_scan_latency(cx, workload="vlp16"):
    """BASELINE config 2 as the ROS node runs it: ONE scan per call through the synchronous C ABI (pageable host buffers,
    stream sync per call).  lm_cluster = 16: one thread-block cluster owns the scan for the whole call (fused one-launch
    kernel, scan2map_fused.cu); "with_submap_build" adds msfl_set_submap per frame, as the reference rebuilds its kd-trees
    every frame.  Measured through the Python binding (~40 us of ctypes marshalling per call on top of the C ABI)."""
    from msf_loam_b200 import Engine, default_params
    traj, scans = cx.raw[workload]
    out = {}
    e0 = Engine(default_params(**OVER), device=cx.local_rank)
    mc, ms, queries, _ = build_case(lambda x, r: e0.extract_features(x, r, None), e0.voxel_grid, workload, traj, scans[:WORKLOADS[workload]["map_scans"] + 1])
    e0.close()
    c0, s0, gt = queries[0]
    init = S.perturb_pose(gt, np.random.default_rng(1))
    for G in (1, 16):
        e1 = Engine(default_params(lm_cluster=G, **OVER), device=cx.local_rank)
        e1.set_submap(mc, ms)
        for _ in range(5):
            e1.scan2map(c0, s0, init, want_stats=False)
        t0 = time.perf_counter()
        for _ in range(40):
            e1.scan2map(c0, s0, init, want_stats=False)
        out[f"lm_cluster_{G}_us_per_scan"] = round((time.perf_counter() - t0) / 40 * 1e6, 1)
        t0 = time.perf_counter()
        for _ in range(20):  # the reference rebuilds the kd-trees every frame (mapping_scan_matcher.cc:66-72)
            e1.set_submap(mc, ms)
            e1.scan2map(c0, s0, init, want_stats=False)
        out[f"lm_cluster_{G}_us_per_scan_with_submap_build"] = round((time.perf_counter() - t0) / 20 * 1e6, 1)
        e1.close()
    # the caller's whole frame (LaserMapping::MatchScan2Map + InsertScan2Map, laser_mapping.cc:258-340) as ONE call:
    # VoxelGrid of the scan's feature clouds, GetSurroundedCloud of both GPU-resident maps, the gate, MatchScan2Map,
    # InsertScan at the refined pose -- the un-down-sampled less-sharp / less-flat clouds in, the pose out
    from msf_loam_b200 import HybridGrid, mapping_frame
    e2 = Engine(default_params(lm_cluster=16), device=cx.local_rank)  # the reference's own schedule (Ceres termination tests on)
    n_map = WORKLOADS[workload]["map_scans"]
    frames = []
    for k in range(min(len(scans), n_map + 12)):
        f = e2.extract_features(scans[k][0], scans[k][1], None)
        frames.append((f["full"][f["idx_less_sharp"]], f["full"][f["idx_less_flat"]], traj[k]))
    gc, gs = HybridGrid(e2, 3.0, 0.2), HybridGrid(e2, 3.0, 0.4)
    rng, ts, errs = np.random.default_rng(3), [], []
    for k, (corner, surf, gt) in enumerate(frames):
        guess = gt if k == 0 else S.perturb_pose(gt, rng)
        t0 = time.perf_counter()
        matched, pose, _ = mapping_frame(e2, gc, gs, corner, surf, guess, want_stats=False)
        ts.append(time.perf_counter() - t0)
        if matched:
            errs.append(S.pose_error(pose, gt))
    warm = ts[len(ts) // 2:]
    out["mapping_frame_us"] = round(float(np.mean(warm)) * 1e6, 1)
    out["mapping_frame"] = {"frames": len(frames), "timed": len(warm), "corner_points": int(frames[-1][0].shape[0]),
                            "surf_points": int(frames[-1][1].shape[0]), "map_points": [gc.size()[0], gs.size()[0]],
                            "max_err_vs_ground_truth_m": float(max(e[0] for e in errs)),
                            "api": "msfl_mapping_frame: one call per frame, every intermediate on the device"}
    gc.close(); gs.close(); e2.close()
    return out


def deskew_batch(cx, workload="vlp16", B=512, steps=5):
    """SURVEY.md 8f row 3 at batch scale: B scans of a replayed log through the IMU-initialised branch of MatchScan2Map
    (Deskew factors, per-point GetDeltaQP) in ONE msfl_scan2map_deskew_batch call; every scan has its own preintegration
    buffers (400 Hz over the scan period), velocity and gravity.  Host buffers -> host poses; pose check vs the oracle."""
    import torch
    from msf_loam_b200 import Engine, default_params
    traj, scans = cx.raw[workload]
    eng = Engine(default_params(**OVER), device=cx.local_rank)
    mc, ms, queries, _ = build_case(lambda x, r: eng.extract_features(x, r, None), eng.voxel_grid, workload, traj, scans)
    eng.set_submap(mc, ms)
    D = len(queries)
    rng = np.random.default_rng(3000 + cx.rank)
    t = np.arange(0.0, 0.105 + 1e-9, 1.0 / 400.0)
    tabs = []
    for b in range(B):
        omega, acc, v0 = rng.normal(scale=0.05, size=3), rng.normal(scale=0.3, size=3), rng.normal(scale=0.05, size=3)
        dq = np.stack([S.rotvec_to_quat(omega * ti) for ti in t])
        dp = np.stack([v0 * ti + 0.5 * acc * ti * ti for ti in t])
        tabs.append((t, dq, dp, tuple(rng.normal(scale=0.2, size=3)), (0.0, 0.0, 9.81)))
    corners, surfs = [queries[b % D][0] for b in range(B)], [queries[b % D][1] for b in range(B)]
    inits = np.stack([S.perturb_pose(queries[b % D][2], rng) for b in range(B)])
    batch = eng.prepare_deskew_batch(corners, surfs, tabs)
    poses = inits.copy()
    eng.scan2map_deskew_prepared(batch, poses)
    x = inits.copy()
    eng.scan2map_deskew_prepared(batch, x)
    torch.cuda.synchronize(cx.dev)
    eng.set_profiling(True)
    eng.get_profile()
    t0 = time.perf_counter()
    for _ in range(steps):
        x[:] = inits
        eng.scan2map_deskew_prepared(batch, x)
    dt = time.perf_counter() - t0
    gpu_ms, _ = eng.get_profile()
    eng.set_profiling(False)
    t1 = time.perf_counter()
    for b in range(8):
        eng.scan2map_deskew(corners[b], surfs[b], *tabs[b], inits[b], want_stats=False)
    dt1 = (time.perf_counter() - t1) / 8
    eng.close()
    rec = {"value": round(B * steps / dt, 1), "unit": "scans/s", "ms_per_step": round(dt / steps * 1e3, 3), "scans_per_step": B,
           "one_scan_per_call_us": round(dt1 * 1e6, 1), "preintegration_samples_per_scan": int(t.shape[0]),
           "gpu_ms_per_step": {"association": round(gpu_ms[0] / steps, 3), "lm_solve": round(gpu_ms[1] / steps, 3)},
           "api": "msfl_scan2map_deskew_batch (mapping_scan_matcher.cc with is_initialized == true, LiDAR part): host clouds "
                  "(one array per cloud, pageable) + preintegration tables -> host poses, one synchronous call per step"}
    if not cx.args.no_cpu and cx.rank == 0:
        import oracle as O
        P = O.default_params(**OVER)
        errs = []
        for b in range(min(4, B)):
            _, x, _, _, _ = O.scan2map_deskew(P, mc, ms, corners[b], surfs[b], *tabs[b], inits[b])
            errs.append(S.pose_error(poses[b], x))
        rec["pose_err_vs_oracle"] = {"max_trans_m": float(max(e[0] for e in errs)), "max_rot_rad": float(max(e[1] for e in errs)),
                                     "scans_checked": len(errs), "tolerance": "1e-4 m / 1e-4 rad"}
    return rec


def chain_raw_to_pose(cx, workload="vlp16", B=256, steps=5):
    """Whole chain through ONE C-ABI call per batch (msfl_register_and_match_batch): B raw scans in the reference's
    PointXYZIRT layout (32 B points, pageable host memory) -> scan registration -> VoxelGrid 0.2 / 0.4 -> scan-to-map,
    every stage batched on the GPU; timed from host buffers to host poses, checked against the oracle's own chain."""
    import torch
    from msf_loam_b200 import Engine, default_params
    traj, scans = cx.raw[workload]
    n_map = WORKLOADS[workload]["map_scans"]
    eng = Engine(default_params(**OVER), device=cx.local_rank)
    mc, ms, queries, n_full = build_case(lambda x, r: eng.extract_features(x, r, None), eng.voxel_grid, workload, traj, scans)
    eng.set_submap(mc, ms)
    D = len(queries)
    rng = np.random.default_rng(2000 + cx.rank)
    raws = [scans[n_map + (i % D)] for i in range(B)]
    # every scan of the batch owns its arrays, as B independent messages would
    batch = eng.prepare_raw_batch([r[0].copy() for r in raws], [r[1].copy() for r in raws])
    inits = np.stack([S.perturb_pose(queries[i % D][2], rng) for i in range(B)])
    poses, counts = eng.register_and_match_batch(batch, inits, want_counts=True)
    for _ in range(2):
        eng.register_and_match_batch(batch, inits)
    torch.cuda.synchronize(cx.dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.register_and_match_batch(batch, inits)
    dt = time.perf_counter() - t0
    # the same replay WITH the odometry of the B - 1 consecutive pairs (msfl_replay_batch): the scans run forwards and
    # backwards along the trajectory so that every pair of the batch is a pair of neighbouring scans
    tri = [(i % (2 * D - 2)) if D > 1 else 0 for i in range(B)]
    tri = [t if t < D else 2 * D - 2 - t for t in tri]
    raws_o = [scans[n_map + t] for t in tri]
    batch_o = eng.prepare_raw_batch([r[0].copy() for r in raws_o], [r[1].copy() for r in raws_o])
    inits_o = np.stack([S.perturb_pose(queries[t][2], rng) for t in tri])
    ident = np.tile(S.pose_identity(), (B, 1))
    od, od_status, poses_o = eng.replay_batch(batch_o, ident, inits_o, compose=False)
    eng.replay_batch(batch_o, ident, inits_o, compose=False)
    torch.cuda.synchronize(cx.dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.replay_batch(batch_o, ident, inits_o, compose=False)
    dt_o = time.perf_counter() - t0
    od_gt = [S.pose_error(od[b], S.pose_mul(S.pose_inv(queries[tri[b - 1]][2]), queries[tri[b]][2])) for b in range(1, B)]
    with_odo = {"value": round(B * steps / dt_o, 1), "unit": "scans/s", "ms_per_step": round(dt_o / steps * 1e3, 3),
                "pairs_per_step": B - 1, "pairs_ok": int((od_status[1:] == 0).sum()),
                "odometry_err_vs_ground_truth": {"max_trans_m": float(max(e[0] for e in od_gt)), "max_rot_rad": float(max(e[1] for e in od_gt))},
                "api": "msfl_replay_batch: the call above plus OdometryScanMatcher::MatchScan2Scan of the B - 1 consecutive pairs "
                       "(identity initial guesses) from the same registration pass; one synchronous call per step"}
    eng.close()
    rec = {"value": round(B * steps / dt, 1), "unit": "scans/s", "ms_per_step": round(dt / steps * 1e3, 3), "scans_per_step": B,
           "points_per_scan": n_full, "queries_per_scan": round(float(np.mean([c["n_corner_queries"] + c["n_surf_queries"] for c in counts])), 1),
           "h2d_bytes_per_step": int(sum(r[0].shape[0] for r in raws) * 18), "d2h_bytes_per_step": B * 56,
           "api": "msfl_register_and_match_batch: raw PointXYZIRT clouds (32 B points, pageable) -> registration -> VoxelGrid -> "
                  "scan-to-map, one launch sequence per stage for the whole batch; one synchronous call per step"}
    if not cx.args.no_cpu and cx.rank == 0:
        import oracle as O
        P = O.default_params(**OVER)
        errs = []
        for i in range(min(4, B)):
            f = O.extract_features(P, raws[i][0], raws[i][1], None)
            qc, qs = O.voxel_grid(f["full"][f["idx_less_sharp"]], 0.2), O.voxel_grid(f["full"][f["idx_less_flat"]], 0.4)
            x, _, _ = O.scan2map(P, mc, ms, qc, qs, inits[i])
            errs.append(S.pose_error(poses[i], x))
        rec["pose_err_vs_oracle"] = {"max_trans_m": float(max(e[0] for e in errs)), "max_rot_rad": float(max(e[1] for e in errs)),
                                     "scans_checked": len(errs), "tolerance": "1e-4 m / 1e-4 rad"}
        errs = []
        fo = [O.extract_features(P, raws_o[i][0], raws_o[i][1], None) for i in range(min(3, B))]
        for b in range(1, len(fo)):
            f0, f1 = fo[b - 1], fo[b]
            _, x, _, _, _ = O.scan2scan(P, f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]],
                                        f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]],
                                        f1["full"][f1["idx_sharp"]], f1["full"][f1["idx_flat"]], S.pose_identity())
            errs.append(S.pose_error(od[b], x))
        if errs:
            with_odo["odometry_err_vs_oracle"] = {"max_trans_m": float(max(e[0] for e in errs)), "max_rot_rad": float(max(e[1] for e in errs)),
                                                  "pairs_checked": len(errs), "tolerance": "1e-4 m / 1e-4 rad"}
    rec["with_odometry"] = with_odo
    return rec


def run_ours(args):
    rank, local_rank, world = dist_env()
    cx = Ctx()
    cx.args, cx.rank, cx.local_rank, cx.world = args, rank, local_rank, world
    extra = not args.no_workloads and not args.only_device
    # raw scans first: numpy ray-casting in a fork pool, before this process owns a CUDA context
    n_workers = max(1, (os.cpu_count() or 1) // world)
    os.environ.setdefault("MSFL_PACK_THREADS", str(max(1, min(16, n_workers))))
    t_gen = time.perf_counter()
    cx.raw = {args.workload: raw_scans(args.workload, args.distinct, n_workers)}
    if extra:
        for wl, nd in (("hdl64", 8), ("os1-128", 4), ("vlp16", 8)):
            if wl not in cx.raw:
                cx.raw[wl] = raw_scans(wl, nd, n_workers)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cx.dev = torch.device("cuda", local_rank)
    numa = numa_bind(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=cx.dev)

    # the CPU legs run on rank 0 at N = 1 only (at N > 1 the other ranks would idle in a barrier meanwhile)
    main = run_workload(cx, args.workload, args.batch, args.distinct, args.steps, args.warmup, full=True,
                        cpu_scans=args.cpu_sample if world == 1 else 64)
    workloads = None
    if extra:
        workloads = {}
        st = max(args.steps, 20)
        # config 3: HDL-64E
        workloads["config3_hdl64"] = run_workload(cx, "hdl64", 512, 8, st, 3, full=False, sampler=False, cpu_scans=8)
        # config 5: OS1-128 vs the ~1 M-point submap, batch sweep per GPU (large scans in small batches: the solve is
        # spread over thread-block clusters so that the LM kernel fills the chip)
        for b, g in ((64, 4), (8, 8), (1, 8)):
            workloads[f"config5_os1-128_batch{b}"] = run_workload(cx, "os1-128", b, 4, st, 3, full=False, lm_cluster=g,
                                                                 sampler=False, e2e=(b == 64), cpu_scans=4 if b == 64 else 0)
        # config 4 as worded: one VLP-16 scan per GPU; every step = new map version indexed on rank 0, broadcast,
        # adopted, one scan solved per rank; ms_per_step is the latency of that whole exchange
        workloads["config4_one_scan_per_gpu"] = run_workload(cx, "vlp16", 1, 8, st * 5, 5, full=False, lm_cluster=8,
                                                             sampler=False, e2e=False, cpu_scans=0, remap=True)
        # the whole chain raw cloud -> pose at batch throughput (rank-local; every rank runs its own batch)
        ch = chain_raw_to_pose(cx)
        if world > 1:
            t = torch.tensor([ch["value"]], dtype=torch.float64, device=cx.dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ch["value"] = round(float(t[0]), 1)
        workloads["chain_raw_to_pose_vlp16"] = ch
        # the IMU-initialised (Deskew) branch at batch scale, rank-local like the chain
        dk = deskew_batch(cx)
        if world > 1:
            t = torch.tensor([dk["value"]], dtype=torch.float64, device=cx.dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dk["value"] = round(float(t[0]), 1)
        workloads["deskew_branch_batch_vlp16"] = dk
    single = single_scan_latency(cx) if (rank == 0 and not args.only_device) else None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    clocks = main.pop("clocks") or {}
    out = {
        "metric": "scans/sec scan-to-map", "value": main["value"], "unit": "scans/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": main["config"], "impl_notes": main["impl_notes"], "e2e": main.get("e2e"),
        "gpu_launches": main["gpu_launches"], "roofline": main["roofline"],
        "cpu_baseline": main.get("cpu_baseline"), "pose_err_vs_oracle": main.get("pose_err_vs_oracle"),
        "single_scan_latency": single,
        "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples", "samples_under_load", "power_w_max")},
        "host": {"numa": numa, "input_generation_s": round(t_gen, 1)},
    }
    if workloads is not None:
        for w in workloads.values():
            w.pop("clocks", None)
        out["workloads"] = workloads
    return out


def cpu_baseline(map_corner, map_surf, qc, c_off, qs, s_off, inits, poses_gpu, n_sample, over):
    """Oracle (kind "port"), one thread, on a bounded sample of the same batch; also the pose check."""
    import oracle as O
    P = O.default_params(**{k: v for k, v in over.items() if k != "lm_cluster"})
    n = n_sample
    co, so = c_off[: n + 1], s_off[: n + 1]
    t0 = time.perf_counter()
    poses = O.scan2map_batch(P, map_corner, map_surf, qc[: co[-1]], co, qs[: so[-1]], so, inits[:n], n_threads=1)
    dt = time.perf_counter() - t0
    errs = [S.pose_error(poses_gpu[i], poses[i]) for i in range(n)]
    pose_err = {"max_trans_m": float(max(e[0] for e in errs)), "max_rot_rad": float(max(e[1] for e in errs)),
                "scans_checked": n, "tolerance": "1e-4 m / 1e-4 rad"}
    cpu = {"value": round(n / dt, 2), "unit": "scans/s", "cores": 1, "kind": "port",
           "sample": f"first {n} scans of the batch, {dt:.1f} s, single thread, kd-trees built once for the batch",
           "host_cpus": os.cpu_count()}
    return cpu, pose_err


# ------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference path on all host threads
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return None
    import oracle as O
    threads = os.cpu_count() or 1
    P = O.default_params(**OVER)
    # the same workload as the CUDA arm: same scans, same submap, same batch assembly (the oracle's extraction and
    # VoxelGrid are bit-identical to the CUDA ones, so `config` is equal key by key)
    traj, scans = raw_scans(args.workload, args.distinct, threads)
    map_corner, map_surf, queries, n_full = build_case(lambda x, r: O.extract_features(P, x, r, None), O.voxel_grid,
                                                       args.workload, traj, scans)
    B = args.batch
    qc, c_off, qs, s_off, inits = assemble_batch(queries, B, seed=1000)
    n_q = int(c_off[-1] + s_off[-1])
    n = min(B, max(threads * 2, 8))  # bounded sample per step: the first n scans of the batch
    co, so = c_off[: n + 1], s_off[: n + 1]
    sample = (qc[: co[-1]], co, qs[: so[-1]], so, inits[:n])
    for _ in range(max(1, min(args.warmup, 2))):
        O.scan2map_batch(P, map_corner, map_surf, *sample, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.scan2map_batch(P, map_corner, map_surf, *sample, n_threads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    # the reference's OWN matcher, compiled unmodified into oracle/_ref (stand-in third-party headers, see
    # oracle/ref_shim.cc), on the same sample and schedule: every call builds its kd-trees like the reference does each
    # frame.  Reported next to the port; the headline of this arm stays the faster of the two (the port).
    ref_compiled = None
    try:
        from oracle import ref as R
        if R.available():
            k = max(1, args.steps // 4)
            R.scan2map_batch(map_corner, map_surf, *sample, n_threads=threads, fixed_attempts=L_ATTEMPTS)
            t0 = time.perf_counter()
            for _ in range(k):
                x_ref = R.scan2map_batch(map_corner, map_surf, *sample, n_threads=threads, fixed_attempts=L_ATTEMPTS)
            dt_ref = time.perf_counter() - t0
            x_port = O.scan2map_batch(P, map_corner, map_surf, *sample, n_threads=threads)
            ref_compiled = {"value": round(n * k / dt_ref, 2), "unit": "scans/s", "cores": threads, "kind": "reference",
                            "steps": k, "max_abs_pose_diff_vs_port": float(np.abs(x_ref - x_port).max()),
                            "what": "MappingScanMatcher::MatchScan2Map of mapping_scan_matcher.cc compiled unmodified "
                                    "(oracle/_ref/libmsfl_ref.so; PCL / Eigen / Ceres stood in, the solver loop is the port's), "
                                    "one call per scan incl. its two kd-tree builds"}
    except Exception as ex:  # the prebuilt library is optional on the GPU box
        ref_compiled = {"unavailable": str(ex)[:120]}
    return {
        "impl": "reference", "metric": "scans/sec scan-to-map", "value": round(value, 2), "unit": "scans/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, B, queries, n_full, n_q, int(map_corner.shape[0]), int(map_surf.shape[0])),
        "cpu_baseline": {"value": round(value, 2), "unit": "scans/s", "cores": threads, "kind": "port",
                         "sample": f"first {n} scans of the {B}-scan batch per step x {args.steps} steps, pthreads over "
                                   "independent scans, kd-trees built once per step; CPU restatement of the reference "
                                   "PCL+Ceres path (libraries not installable offline); bit-equal to the reference's "
                                   "own matcher sources compiled against stand-in headers (reference_compiled)",
                         "reference_compiled": ref_compiled},
        "e2e": {"value": round(value, 2), "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    args = parse_args()
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
