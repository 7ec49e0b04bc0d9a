#!/bin/bash
# Experiment helper: kernel-only durations (ncu launch list) of the merge kernels for library variants.
# usage: scripts/exp_fuse_ncu.sh name...   → per variant: mean µs per kernel name over the last 60 launches of a 40-step builder run
for l in "$@"; do
  DM_B200_LIB=build/exp/lib_$l.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"fuse_" -c 400 \
    --log-file gpurun_out/exp_fuse_$l.csv python bench.py --workload builder --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2 > /dev/null 2>&1
  python - $l <<'PY'
import csv, io, sys, collections
l = sys.argv[1]
rows = list(csv.DictReader(io.StringIO("".join(x for x in open(f"gpurun_out/exp_fuse_{l}.csv") if x.startswith('"')))))
rows = rows[-75:]
by = collections.defaultdict(list)
for r in rows:
  by[r["Kernel Name"].split("(")[0] + " " + r["Grid Size"]].append(float(r["Metric Value"]) / 1e3)
print(l, {k: round(sum(v) / len(v), 1) for k, v in sorted(by.items())}, "sum/step ≈", round(sum(sum(v) for v in by.values()) / 15 * 1.0, 1))
PY
done
