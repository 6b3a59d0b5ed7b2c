"""Dev helper: per-source-line hot spots (samples, executed instructions) for one kernel from an .ncu-rep.
usage: python tools/ncu_source_hot.py rep.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find the first header row
agg = {}
hdr = None
cur_file = ""
done_kernels = 0
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if r and r[0] == "Function Name":
        done_kernels += 1
        if done_kernels > 1:
            break
    if r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); i_t = hdr.index("Thread Instructions Executed")
        continue
    if hdr and r and r[0].isdigit():
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0, 0])
        f=lambda v: int(v) if v.strip().lstrip('-').isdigit() else 0
        a[0] += f(r[i_s]); a[1] += f(r[i_i]); a[2] += f(r[i_t])
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot_s:5.1f}% smp {100*a[1]/tot_i:5.1f}% inst  lanes {a[2]/max(a[1],1):4.1f}  {f}:{ln}  {src}")
