"""python profiles/ingest.py rNN [dir] -- turns <dir>/{launches.csv,prof_main.ncu-rep,bench_line.json} (default dir:
gpurun_out) into profiles/rNN_launches.csv, rNN_kernels.md, rNN_bench.json and refreshes profiles/traffic.json; further
captures <dir>/prof_<name>.ncu-rep are appended to rNN_kernels.md."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, sys.argv[2] if len(sys.argv) > 2 else "gpurun_out")
prof = os.path.join(ROOT, "profiles")

shutil.copy(os.path.join(out, "launches.csv"), os.path.join(prof, f"{tag}_launches.csv"))
shutil.copy(os.path.join(out, "bench_line.json"), os.path.join(prof, f"{tag}_bench.json"))

# launch shares
rows = list(csv.reader(open(os.path.join(out, "launches.csv"))))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[hi + 1:]:
    if len(r) > mv:
        agg.setdefault(r[kn].split("(")[0], []).append(float(r[mv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
lines = [f"# {tag}: launches of `python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --scenes 1` (ncu "
         "gpu__time_duration, cold cache, serialised)\n",
         "| kernel | launches | avg us | share |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f} % |")

# full capture
summary = subprocess.run([sys.executable, os.path.join(prof, "summarize.py"), os.path.join(out, "prof_main.ncu-rep")],
                         capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", os.path.join(out, "prof_main.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u = rr[0], rr[1]
traffic = {}
for r in rr[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
    def gb(key):
        i = h.index(key)
        val = float(r[i])
        return val * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[i]]
    traffic[name] = int(gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"))
tj = os.path.join(prof, "traffic.json")
data = json.load(open(tj)) if os.path.exists(tj) else {}
data["cfg2"] = dict(traffic, _source=f"ncu --set full, {tag}: profiles/{tag}_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)")
json.dump(data, open(tj, "w"), indent=1)

bench = json.loads([l for l in open(os.path.join(out, "bench_line.json")) if l.startswith("{")][-1])
# the profiled command runs ONE scene (seed 0); the bench line rotates four: note the rows of the captured fill
one = os.path.join(out, "bench_one_scene.json")
if os.path.exists(one):
    b1 = json.loads([l for l in open(one) if l.startswith("{")][-1])
    data["cfg2"]["_fill_rows"] = b1["rows_per_scene"]["max"]
    json.dump(data, open(tj, "w"), indent=1)
lines += ["", f"Un-profiled bench line of the same build: {bench['ms_per_step']:.3f} ms/step, "
              f"stage_ms = {bench['stage_ms']}, clocks = {bench['clocks']}", "", summary]
import glob
for extra in sorted(glob.glob(os.path.join(out, "prof_*.ncu-rep"))):
    if extra.endswith("prof_main.ncu-rep"):
        continue
    lines += ["", subprocess.run([sys.executable, os.path.join(prof, "summarize.py"), extra], capture_output=True, text=True).stdout]
open(os.path.join(prof, f"{tag}_kernels.md"), "w").write("\n".join(lines))
print("\n".join(lines[:12]))
print(traffic)
