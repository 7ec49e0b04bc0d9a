#!/bin/bash
for lib in mc32k mc160k; do
  DM_B200_LIB=build/exp/lib_$lib.so ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"fuse_|hmap_" -s 1300 -c 48 --log-file gpurun_out/mc_$lib.csv python scripts/time_builder.py builder -1 > /dev/null 2>&1
  python - $lib <<'PY'
import csv, sys, collections
d = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/mc_{d}.csv")) if len(r) > 5 and r[0].isdigit()]
per = collections.defaultdict(list)
for r in rows:
  per[r[4].split("(")[0][:40]].append(float(r[-1].replace(",", "")))
for k, v in per.items():
  print(f"{d} {k:28s} n={len(v):3d} mean {sum(v)/len(v)/1e3:8.1f} us  last {v[-1]/1e3:8.1f}")
PY
  DM_B200_LIB=build/exp/lib_$lib.so python scripts/time_builder.py builder -1
done
