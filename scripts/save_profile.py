#!/usr/bin/env python
"""Copies the judged evidence of one gpurun call from gpurun_out/ (scratch) into profiles/ (tracked):
  profiles/<tag>_launches.csv       ncu launch list (gpu__time_duration per launch) of the bench command
  profiles/<tag>_proj_full.txt      headline metrics / opcode mix / stall summary of the ncu --set full capture
  profiles/<tag>_bench_<scene>.json the bench lines of the same call
  profiles/traffic.json             dram bytes per launch of the dominant kernel (read by bench.py)
usage: scripts/save_profile.py <run-tag> [<profile-tag>]      e.g. save_profile.py r1a r01
"""
import csv
import glob
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

def short(name):
  return name.split("(")[0].replace("void ", "")


def save_launches(src, dst, cmd):
  if not os.path.exists(src):
    return
  lines = [l for l in open(src) if l.startswith('"')]
  rows = list(csv.DictReader(io.StringIO("".join(lines))))
  with open(dst, "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ... {cmd}   (cold-cache, serialised: compare shares, not absolutes)\n")
    f.write("id,kernel,grid,block,stream,duration_ns\n")
    for r in rows:
      f.write(f'{r["ID"]},{short(r["Kernel Name"])},"{r["Grid Size"]}","{r["Block Size"]}",{r["Stream"]},{r["Metric Value"]}\n')
  tot = sum(float(r["Metric Value"]) for r in rows)
  by = {}
  for r in rows:
    by[short(r["Kernel Name"])] = by.get(short(r["Kernel Name"]), 0.0) + float(r["Metric Value"])
  print(os.path.basename(dst), "launch shares:", {k: f"{100 * v / tot:.1f}%" for k, v in by.items()}, f"{len(rows)} launches")


save_launches(os.path.join(G, f"launches_{run}.csv"), os.path.join(P, f"{tag}_launches.csv"),
              "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2")
save_launches(os.path.join(G, f"launches_labels_{run}.csv"), os.path.join(P, f"{tag}_launches_labels.csv"),
              "python bench.py --workload proj_labels --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 2")
save_launches(os.path.join(G, f"launches_flow_{run}.csv"), os.path.join(P, f"{tag}_launches_flow.csv"),
              "python bench.py --workload flow --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2")
save_launches(os.path.join(G, f"launches_builder_{run}.csv"), os.path.join(P, f"{tag}_launches_builder.csv"),
              "python bench.py --workload builder --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2")


def raw_rows(rep):
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  H, units = rows[0], rows[1]

  def val(r, name):
    v, u = float(r[H.index(name)]), units[H.index(name)]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
  return [dict(kernel=short(r[H.index("Kernel Name")]), rd=val(r, "dram__bytes_read.sum"),
               wr=val(r, "dram__bytes_write.sum"), us=float(r[H.index("gpu__time_duration.sum")])) for r in rows[2:]]


traffic_path = os.path.join(P, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for what, pattern, n_kernels in (("proj", "proj_ws", 1), ("labels", "proj_lbl", 1), ("flow", "flow_", 1),
                                 ("fuse", "fuse_|changed_", 6)):
  rep = os.path.join(G, f"prof_{what}_{run}.ncu-rep")
  if not os.path.exists(rep):
    continue
  with open(os.path.join(P, f"{tag}_{what}_full.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{pattern}  (run {run})\n")
    for i in range(n_kernels):
      out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, str(i), "30"],
                           capture_output=True, text=True).stdout
      f.write(out + "\n")
  rows = raw_rows(rep)
  if what == "proj":
    r = rows[0]
    traffic.update({"kernel": r["kernel"], "run": run, "dram_bytes_read": r["rd"], "dram_bytes_write": r["wr"],
                    "dram_bytes_per_step": r["rd"] + r["wr"], "duration_us_under_ncu": r["us"]})
  elif what == "labels":
    r = rows[0]
    traffic["proj_labels"] = {"kernel": r["kernel"], "run": run, "dram_bytes_read": r["rd"], "dram_bytes_write": r["wr"],
                              "dram_bytes_per_step": r["rd"] + r["wr"], "duration_us_under_ncu": r["us"]}
  elif what == "flow":
    r = rows[0]
    traffic["flow"] = {"kernel": r["kernel"], "run": run, "dram_bytes_per_step": r["rd"] + r["wr"],
                       "duration_us_under_ncu": r["us"]}
  else:  # one merge = bbox_init + bbox + fill + scatter: sum the consecutive kernels of one step
    names = [r["kernel"] for r in rows]
    # a merge starts with the bbox_init of pass 1 (followed by a bbox kernel; the second bbox_init of a merge
    # arms the tracked box of pass 2 and is followed by a scatter kernel)
    inits = [i for i, k in enumerate(names[:-1]) if k == "fuse_bbox_init" and names[i + 1] == "fuse_bbox_kernel"]
    if len(inits) >= 2:  # one complete merge: everything between two pass-1 bbox_init launches
      step = rows[inits[0]:inits[1]]
      traffic["builder"] = {"kernels": [r["kernel"] for r in step], "run": run,
                            "dram_bytes_per_step": sum(r["rd"] + r["wr"] for r in step),
                            "duration_us_under_ncu": sum(r["us"] for r in step)}
  print(what, "traffic ok")
json.dump(traffic, open(traffic_path, "w"), indent=1)

for name in ("default", "room", "room20", "iid", "labels", "labels20", "labels_iid", "flow", "builder", "builder_fixed",
             "proj5", "n2", "ref"):
  b = os.path.join(G, f"bench_{name}_{run}.json")
  if os.path.exists(b) and os.path.getsize(b):
    shutil.copy(b, os.path.join(P, f"{tag}_bench_{name}.json"))

# compute-sanitizer logs of the small parity cases, host backtraces dropped
for tool in ("memcheck", "racecheck"):
  src = os.path.join(G, f"sanitizer_{tool}_{run}.log")
  if os.path.exists(src):
    keep = [l for l in open(src, errors="replace") if "Host Frame" not in l and "frame #" not in l]
    with open(os.path.join(P, f"{tag}_sanitizer_{tool}.log"), "w") as f:
      f.write(f"# compute-sanitizer --tool {tool} --error-exitcode 9 python -m pytest <7 small parity cases, see scripts/r02_check.sh>  (run {run}; host frames dropped)\n")
      f.writelines(keep[:400])
