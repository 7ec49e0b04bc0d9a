#!/bin/bash
# Kernel experiments on the label kernel (under gpurun): builds of dm_labels.cu with other schedule constants /
# ablation macros (built beforehand with scripts/exp_build.sh into build/exp/), each timed with the proj_labels workload.
mkdir -p gpurun_out
for lib in build/exp/lib_*.so; do
  name=$(basename $lib .so)
  for scene in room iid; do
    DM_B200_LIB=$PWD/$lib timeout 200 python bench.py --workload proj_labels --scene $scene --steps 100 --no-cpu-baseline --e2e-steps 0 > gpurun_out/exp_${name}_${scene}.json 2> gpurun_out/exp_${name}_${scene}.err
    python - $name $scene <<'PY'
import json, sys
try:
  d = json.load(open(f"gpurun_out/exp_{sys.argv[1]}_{sys.argv[2]}.json"))
  print(f"{sys.argv[1]:28s} {sys.argv[2]:5s} ms/step={d['ms_per_step']:.4f} value={d['value']:.0f}")
except Exception as e:
  print(sys.argv[1], sys.argv[2], "failed", e)
PY
  done
done
