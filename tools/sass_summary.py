"""Per-kernel SASS evidence for profiles/: instruction mnemonics that prove what the kernels use (bulk async copies,
mbarriers, cluster barriers, fp64 pipe, no tensor-core MMA) + registers / spills from the embedded resource usage.
usage: python tools/sass_summary.py [libmsfl.so] > profiles/rN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "msf_loam_b200/libmsfl.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WATCH = ["UBLKCP", "SYNCS", "UCGABAR", "DFMA", "DMUL", "DADD", "MUFU", "FFMA", "LDG", "LDS", "STS", "ATOM", "RED", "SHFL", "LDL", "STL",
         "HMMA", "UTCMMA", "UTCHMMA", "QGMMA", "IMMA"]
kern = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        kern[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                kern[cur][w] += 1
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups())
print(f"SASS summary of {lib} (cuobjdump -sass / -res-usage; sm_100a).  Columns: instruction counts by mnemonic prefix.")
print("UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier ops, UCGABAR = cluster barrier, D* = fp64 pipe,")
print("HMMA/UTC*MMA/QGMMA/IMMA = tensor-core MMA (expected 0: no contraction on this path).\n")
for k, c in kern.items():
    d = demangle(k)
    if "msfl::" not in d:
        continue
    short = re.sub(r"\(.*", "", d).replace("msfl::", "").replace("void ", "")
    u = usage.get(k, (0, 0, 0, 0))
    cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{short:<44} regs={u[0]:<3} stack={u[1]:<4} smem={u[2]:<6} instr={c['_total']:<6} {cols}")
tc = sum(c[w] for c in kern.values() for w in ("HMMA", "UTCMMA", "UTCHMMA", "QGMMA", "IMMA"))
print(f"\ntensor-core MMA instructions in the whole library: {tc}")
