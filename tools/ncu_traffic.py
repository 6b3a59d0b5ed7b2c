"""Per-stage DRAM traffic of one bench step from an .ncu-rep -> an entry of profiles/roofline_traffic.json.
usage: python tools/ncu_traffic.py rep.ncu-rep workload:batch "source text" [profiles/roofline_traffic.json]
Stage map (bench.py STAGE_NAMES): 0 = k_knn5_fit / k_knn5, 1 = k_lm_solve, 2 = k_transform_keys + k_scatter_perm,
3 = k_fit + k_fit_qr_list.  Each kernel is averaged over its captured launches."""
import csv
import json
import subprocess
import sys

rep, key, source = sys.argv[1], sys.argv[2], sys.argv[3]
path = sys.argv[4] if len(sys.argv) > 4 else "profiles/roofline_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    f = float(v.replace(",", "") or 0)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


per_kernel = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
        to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    short = name.split("(")[0].split("<")[0].replace("void ", "").strip()
    if short == "k_fit":
        short = "k_fit_corner" if "<0, 1, 1, 0>" in name or "(bool)0, (bool)1, (bool)1, (int)0" in name else "k_fit_other"
    per_kernel.setdefault(short, []).append(b)
avg = {k: sum(v) / len(v) for k, v in per_kernel.items()}
stage = {"0": avg.get("k_knn5_fit", 0) + avg.get("k_knn5", 0), "1": avg.get("k_lm_solve", 0),
         "2": avg.get("k_transform_keys", 0) + avg.get("k_scatter_perm", 0),
         "3": avg.get("k_fit_corner", 0) + avg.get("k_fit_other", 0) + avg.get("k_fit_qr_list", 0)}
entry = {k: int(v) for k, v in stage.items()}
entry["source"] = source
try:
    t = json.load(open(path))
except Exception:
    t = {}
t[key] = entry
json.dump(t, open(path, "w"), indent=1)
print(key, entry, {k: len(v) for k, v in per_kernel.items()})
