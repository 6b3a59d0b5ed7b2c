"""Developer micro-benchmark (NOT bench.py): device-resident scan-to-map batch timing using the
conftest case (inputs built with the oracle, so this lives under tests/)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from conftest import make_map_case
from msf_loam_b200 import Engine, default_params
from msf_loam_b200 import synth as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
case = make_map_case()
qs = case["queries"]
rng = np.random.default_rng(0)
corners, surfs, inits = [], [], []
for b in range(B):
    q = qs[b % len(qs)]
    corners.append(q["corner"]); surfs.append(q["surf"]); inits.append(S.perturb_pose(q["gt"], rng))
c_off = np.concatenate([[0], np.cumsum([c.shape[0] for c in corners])]).astype(np.int32)
s_off = np.concatenate([[0], np.cumsum([c.shape[0] for c in surfs])]).astype(np.int32)
dev = torch.device("cuda:0")
d_c = torch.from_numpy(np.concatenate(corners)).to(dev)
d_s = torch.from_numpy(np.concatenate(surfs)).to(dev)
d_co = torch.from_numpy(c_off).to(dev); d_so = torch.from_numpy(s_off).to(dev)
poses0 = torch.from_numpy(np.stack(inits)).to(dev)
over = {"early_exit": 0, "max_num_iterations": 5}
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    e = Engine(default_params(**over), stream=stream.cuda_stream)
    e.set_submap(case["map_corner"], case["map_surf"])
    d_p = poses0.clone()
    def run():
        d_p.copy_(poses0)
        e.scan2map_batch_device(B, d_c.data_ptr(), d_co.data_ptr(), int(c_off[-1]), d_s.data_ptr(), d_so.data_ptr(), int(s_off[-1]), d_p.data_ptr())
    for _ in range(3): run()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps): run()
    ev1.record(stream)
    torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / steps
print(f"B={B} N/scan={(c_off[-1]+s_off[-1])/B:.0f} ms/step={ms:.3f} scans/s={B/ms*1e3:.0f}")
x = d_p.cpu().numpy()
errs = [S.pose_error(x[b], qs[b % len(qs)]["gt"]) for b in range(min(B, 8))]
print("pose err vs GT (first 8):", [f"{a:.4f}/{b:.5f}" for a, b in errs])
