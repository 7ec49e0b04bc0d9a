#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, bench on both scenes, ncu launch list.
# Usage: scripts/gpu_check.sh <tag> [full]    — outputs land in gpurun_out/
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for scene in room iid; do
  timeout 300 python bench.py --scene $scene $([ $scene = iid ] && echo --no-cpu-baseline) > gpurun_out/bench_${scene}_${TAG}.json 2> gpurun_out/bench_${scene}_${TAG}.err
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/bench_${scene}_${TAG}.json"))
  print("${scene}", "value=%.0f maps/s ms/step=%.3f frac=%.3f e2e=%.0f launches=%d clocks=%s cpu=%s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"], (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
  print("bench ${scene} failed:", e); print(open("gpurun_out/bench_${scene}_${TAG}.err").read()[-2000:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"proj_|resolve_|flow_|fuse_" -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches_${TAG}.log 2>&1
grep -c proj_ gpurun_out/launches_${TAG}.csv
if [ "$2" = full ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"proj_ws" -s 4 -c 1 -f -o gpurun_out/prof_proj_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_${TAG}.log
fi
