"""dev: timeline of the pipelined host-buffer path (submit/wait), per-step host timestamps."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from msf_loam_b200 import Engine, default_params
B = 2048
over = {"early_exit": 0, "max_num_iterations": 5, "num_outer": 2}
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    eng = Engine(default_params(**over), device=0, stream=stream.cuda_stream)
    mc, ms, queries, n_full = bench.build_case_gpu(eng, "vlp16", 32, 5)
    qc, c_off, qs, s_off, inits = bench.assemble_batch(queries, B, seed=1000)
    eng.set_submap(mc, ms)
    h_qc = torch.from_numpy(qc).pin_memory(); h_qs = torch.from_numpy(qs).pin_memory()
    hc, hs = h_qc.numpy(), h_qs.numpy()
    prepared = eng.prepare_batch([hc[c_off[i]:c_off[i + 1]] for i in range(B)], [hs[s_off[i]:s_off[i + 1]] for i in range(B)])
    out = [np.zeros_like(inits), np.zeros_like(inits)]
    for depth in (1, 2):
        for rep in range(2):
            torch.cuda.synchronize()
            K = 16
            ts = []
            t0 = time.perf_counter()
            if depth == 1:
                for i in range(K):
                    tk = eng.scan2map_submit(prepared, inits); a = time.perf_counter(); eng.scan2map_wait(tk, out[0]); ts.append((a - t0, time.perf_counter() - t0))
            else:
                tk = eng.scan2map_submit(prepared, inits)
                for i in range(1, K):
                    tk2 = eng.scan2map_submit(prepared, inits); a = time.perf_counter()
                    eng.scan2map_wait(tk, out[i & 1]); ts.append((a - t0, time.perf_counter() - t0)); tk = tk2
                eng.scan2map_wait(tk, out[0]); ts.append((0, time.perf_counter() - t0))
            tot = time.perf_counter() - t0
            print(f"depth {depth}: {K} steps {tot*1e3:.2f} ms -> {B*K/tot:.0f} scans/s; submit-return / wait-return (ms):",
                  " ".join(f"{a*1e3:.1f}/{b*1e3:.1f}" for a, b in ts[:8]))
    # device-resident reference
    import ctypes
    d_qc, d_qs = torch.from_numpy(qc).cuda(), torch.from_numpy(qs).cuda()
    d_co, d_so = torch.from_numpy(c_off).cuda(), torch.from_numpy(s_off).cuda()
    d_p0 = torch.from_numpy(inits).cuda(); d_p = d_p0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        d_p.copy_(d_p0)
        eng.scan2map_batch_device(B, d_qc.data_ptr(), d_co.data_ptr(), int(c_off[-1]), d_qs.data_ptr(), d_so.data_ptr(), int(s_off[-1]), d_p.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    print("device-resident ms/step", e0.elapsed_time(e1) / 10)
