"""dev: H2D bandwidth from pinned memory vs CPU affinity (NUMA placement of the pinned buffer)."""
import os, time, subprocess
import torch
print(subprocess.run("nvidia-smi topo -m; lscpu | grep -i -E 'numa|socket|model name'", shell=True, capture_output=True, text=True).stdout)
n = 160 * 1024 * 1024
def bw(tag):
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(tag, "H2D GB/s", 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
ncpu = os.cpu_count()
bw("default affinity %s" % sorted(os.sched_getaffinity(0)))
for lo in range(0, ncpu, max(1, ncpu // 4)):
    try:
        os.sched_setaffinity(0, set(range(lo, min(ncpu, lo + max(1, ncpu // 4)))))
        bw(f"cpus {lo}-{lo + ncpu // 4 - 1}")
    except Exception as ex:
        print("affinity", lo, ex)
