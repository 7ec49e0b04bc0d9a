#!/usr/bin/env python
"""Copies the judged evidence of one gpurun call from gpurun_out/ (scratch) into profiles/ (tracked):
  profiles/<tag>_launches.csv       ncu launch list (gpu__time_duration per launch) of the bench command
  profiles/<tag>_proj_full.txt      headline metrics / opcode mix / stall summary of the ncu --set full capture
  profiles/<tag>_bench_<scene>.json the bench lines of the same call
  profiles/traffic.json             dram bytes per launch of the dominant kernel (read by bench.py)
usage: scripts/save_profile.py <run-tag> [<profile-tag>]      e.g. save_profile.py r1a r01
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
run = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else run
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# launch list: keep kernel name (short), grid, block, duration
src = os.path.join(G, f"launches_{run}.csv")
if os.path.exists(src):
  lines = [l for l in open(src) if l.startswith('"')]
  rows = list(csv.DictReader(io.StringIO("".join(lines))))
  with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 3 --warmup 3 "
            "--no-cpu-baseline --e2e-steps 2   (cold-cache, serialised: compare shares, not absolutes)\n")
    f.write("id,kernel,grid,block,stream,duration_ns\n")
    for r in rows:
      name = r["Kernel Name"].split("(")[0].replace("void ", "")
      f.write(f'{r["ID"]},{name},"{r["Grid Size"]}","{r["Block Size"]}",{r["Stream"]},{r["Metric Value"]}\n')
  tot = sum(float(r["Metric Value"]) for r in rows)
  by = {}
  for r in rows:
    n = r["Kernel Name"].split("(")[0].replace("void ", "")
    by[n] = by.get(n, 0.0) + float(r["Metric Value"])
  print("launch shares:", {k: f"{100 * v / tot:.1f}%" for k, v in by.items()}, f"{len(rows)} launches")

rep = os.path.join(G, f"prof_proj_{run}.ncu-rep")
if os.path.exists(rep):
  out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, "0", "40"],
                       capture_output=True, text=True).stdout
  with open(os.path.join(P, f"{tag}_proj_full.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:proj_ws -s 4 -c 1  (run {run})\n")
    f.write(out)
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  H, units, r = rows[0], rows[1], rows[2]

  def val(name):
    v, u = float(r[H.index(name)]), units[H.index(name)]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
  rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
  json.dump({"kernel": r[H.index("Kernel Name")].split("(")[0], "run": run,
             "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_step": rd + wr,
             "duration_us_under_ncu": float(r[H.index("gpu__time_duration.sum")])},
            open(os.path.join(P, "traffic.json"), "w"), indent=1)
  print("traffic:", (rd + wr) / 1e9, "GB per launch")

for scene in ("room", "iid"):
  b = os.path.join(G, f"bench_{scene}_{run}.json")
  if os.path.exists(b) and os.path.getsize(b):
    shutil.copy(b, os.path.join(P, f"{tag}_bench_{scene}.json"))
