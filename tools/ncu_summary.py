"""Dev helper: condensed text summary of an .ncu-rep (per kernel: time, regs, occupancy, IPC, lanes,
DRAM bytes, fp64 pipe, warp-stall breakdown).  usage: python tools/ncu_summary.py rep.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
basic = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
         "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
         "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
         "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
         "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
for r in rows[2:]:
    print("=" * 100)
    print(r[col["Kernel Name"]][:150])
    for b in basic:
        if b in col:
            print(f"  {b:70s} {r[col[b]]:>18s} {units[col[b]]}")
    st = [(h, float(r[i] or 0)) for h, i in col.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    tot = sum(v for _, v in st)
    print("  warp stall reasons (warps stalled per issue-active cycle; share):")
    for h, v in sorted(st, key=lambda x: -x[1])[:8]:
        print(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}  {100*v/tot:5.1f}%")
