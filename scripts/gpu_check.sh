#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, bench lines of every workload, ncu launch lists.
# Usage: scripts/gpu_check.sh <tag> [full] [notests]    — outputs land in gpurun_out/
TAG=${1:-rXX}
mkdir -p gpurun_out
if [ "$3" != notests ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
fi
show() {  # show <name> <json>
  python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1:3]
try:
  d = json.load(open(path))
  r = d["roofline"]
  print(name, "value=%.0f %s ms/step=%.3f frac=%.3f (kernel %.3f ms) e2e=%.0f launches=%d clocks=%s cpu=%s" % (
    d["value"], d["unit"], d["ms_per_step"], r["frac"], r.get("kernel_ms_per_step", d["ms_per_step"]), d["e2e"]["value"],
    d["gpu_launches"], d["clocks"], (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
  print("bench", name, "failed:", e)
  print(open(path.replace(".json", ".err")).read()[-2500:])
PY
}
for scene in room iid; do
  timeout 300 python bench.py --scene $scene $([ $scene = iid ] && echo --no-cpu-baseline) > gpurun_out/bench_${scene}_${TAG}.json 2> gpurun_out/bench_${scene}_${TAG}.err
  show $scene gpurun_out/bench_${scene}_${TAG}.json
done
for wl in ${WORKLOADS:-flow builder builder_fixed proj5}; do
  steps=200; [ $wl = proj5 ] && steps=30
  timeout 600 python bench.py --workload $wl --steps $steps > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  show $wl gpurun_out/bench_${wl}_${TAG}.json
done
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU -k regex:"proj_|resolve_|flow_|fuse_" -c 60 --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches_${TAG}.log 2>&1
grep -c proj_ gpurun_out/launches_${TAG}.csv
timeout 300 $NCU -k regex:"flow_" -c 30 --log-file gpurun_out/launches_flow_${TAG}.csv python bench.py --workload flow --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches_flow_${TAG}.log 2>&1
timeout 600 $NCU -k regex:"proj_|resolve_|fuse_|changed_" -c 400 --log-file gpurun_out/launches_builder_${TAG}.csv python bench.py --workload builder --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches_builder_${TAG}.log 2>&1
if [ "$2" = full ]; then
  FULL="ncu --set full --clock-control none --import-source on -f"
  timeout 600 $FULL -k regex:"proj_ws" -s 4 -c 1 -o gpurun_out/prof_proj_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_${TAG}.log
  timeout 600 $FULL -k regex:"flow_" -s 4 -c 1 -o gpurun_out/prof_flow_${TAG} python bench.py --workload flow --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full_flow_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_flow_${TAG}.log
  # the merge kernels of builder step 30 (world map fully grown by then)
  timeout 600 $FULL -k regex:"fuse_|changed_" -s 150 -c 13 -o gpurun_out/prof_fuse_${TAG} python bench.py --workload builder --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full_fuse_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_fuse_${TAG}.log
fi
