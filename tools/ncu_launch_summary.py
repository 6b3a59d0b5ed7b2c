"""Dev helper: per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV) for the launches of the
device-timed steps of bench.py: the full-batch steps (largest k_transform_keys grid) number warmup .. warmup+steps-1.
usage: python tools/ncu_launch_summary.py launches.csv [steps=3] [outer=2] [warmup=3]"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 14 and r[0].isdigit()]
names = [r[4] for r in rows]
t_ns = [float(r[14]) for r in rows]
warmup = int(sys.argv[4]) if len(sys.argv) > 4 else 3


def grid(r):
    return int(r[8].strip("()").split(",")[0])


gmax = max(grid(r) for r in rows if "k_transform_keys" in r[4])
starts = [i for i, r in enumerate(rows) if "k_transform_keys" in r[4] and grid(r) == gmax]
first = starts[warmup * outer]
all_tk = [i for i, r in enumerate(rows) if "k_transform_keys" in r[4]]
after = [i for i in all_tk if i > starts[(warmup + steps) * outer - 1]]
last = after[0] if after else len(rows)
# the step ends with its last k_lm_solve; anything after it (d_p reset copy of the next step) is not ours
while "k_lm_solve" not in names[last - 1]:
    last -= 1
tot = OrderedDict()
for n, t in zip(names[first:last], t_ns[first:last]):
    key = n.split("(")[0].replace("void ", "").replace("msfl::", "").replace("cub::CUB_300001_SM_1000::", "")[:70]
    a = tot.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
total = sum(v[1] for v in tot.values())
print(f"ncu launch list (gpu__time_duration.sum, --clock-control none), the {steps} device-timed steps: "
      f"{sum(v[0] for v in tot.values())} launches, {total / 1e6 / steps:.3f} ms per step (serialised, cold cache)")
print(f"{'kernel':72s}{'launches':>9s}{'ms/step':>10s}{'share':>8s}")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s}{c:9d}{t / 1e6 / steps:10.4f}{100 * t / total:7.1f}%")
