#!/bin/bash
# Round-2 GPU check (under gpurun): parity tests, bench lines (float + label paths), ncu launch list + full capture of
# the label kernel, compute-sanitizer memcheck / racecheck of small parity cases.
# Usage: scripts/r02_check.sh <tag> [notests] [nosan]       — outputs land in gpurun_out/
TAG=${1:-r02x}
export TAG
mkdir -p gpurun_out
if [ "$2" != notests ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; tail -15 gpurun_out/pytest_${TAG}.log | cut -c1-300
fi
show() {
  python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1:3]
try:
  d = json.load(open(path))
  r = d["roofline"]
  print(name, "value=%.0f %s ms/step=%.4f frac=%.3f e2e=%s launches=%d clocks=%s cpu=%s" % (
    d["value"], d["unit"], d["ms_per_step"], r["frac"], d["e2e"]["value"],
    d["gpu_launches"], d["clocks"], (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
  print("bench", name, "failed:", e)
  print(open(path.replace(".json", ".err")).read()[-2500:])
PY
}
run() {  # run <name> <bench args...>
  local name=$1; shift
  timeout 400 python bench.py "$@" > gpurun_out/bench_${name}_${TAG}.json 2> gpurun_out/bench_${name}_${TAG}.err
  show $name gpurun_out/bench_${name}_${TAG}.json
}
run default
python - "$TAG" <<'PY'
import json, sys
try:
  d = json.load(open(f"gpurun_out/bench_default_{sys.argv[1]}.json"))
  print("  e2e", d["e2e"]["value"], "e2e_float32", (d.get("e2e_float32") or {}).get("value"), "pcie", d.get("pcie"))
  for k, v in (d.get("extra") or {}).items():
    print("  extra", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a not in ("workload",)})
except Exception as e:
  print("default line unreadable:", e)
PY
run room20 --steps 20 --warmup 5 --no-cpu-baseline --no-extra --e2e-steps 2
run iid --scene iid --no-cpu-baseline --no-extra --e2e-steps 2
run labels20 --workload proj_labels --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2
run labels_iid --workload proj_labels --scene iid --no-cpu-baseline --e2e-steps 2
for wl in ${WORKLOADS}; do
  run $wl --workload $wl --no-cpu-baseline
done
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU -k regex:"proj_|resolve_" -c 40 --log-file gpurun_out/launches_labels_${TAG}.csv python bench.py --workload proj_labels --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches_labels_${TAG}.log 2>&1
timeout 300 $NCU -k regex:"proj_|resolve_" -c 40 --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 2 > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 600 $NCU -k regex:"proj_|resolve_|fuse_|changed_" -s 1200 -c 200 --log-file gpurun_out/launches_builder_${TAG}.csv python bench.py --workload builder --steps 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches_builder_${TAG}.log 2>&1
timeout 300 $NCU -k regex:"flow_" -c 30 --log-file gpurun_out/launches_flow_${TAG}.csv python bench.py --workload flow --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches_flow_${TAG}.log 2>&1
FULL="ncu --set full --clock-control none --import-source on -f"
# the merge kernels + the height-map projection at the end of the config-4 walk
timeout 600 $FULL -k regex:"fuse_scatter|fuse_fill|fuse_bbox_kernel|hmap_" -s 1500 -c 7 -o gpurun_out/prof_fuse_${TAG} python bench.py --workload builder --steps 100 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full_fuse_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_fuse_${TAG}.log
# the float kernel on 2-D tiles (tensor-map TMA): optional layout, for the record
timeout 600 $FULL -k regex:"proj_ws" -s 5 -c 1 -o gpurun_out/prof_proj2d_${TAG} python scripts/time_proj.py --rows 4 --scene room --steps 3 > gpurun_out/ncu_full_proj2d_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_proj2d_${TAG}.log
timeout 600 $FULL -k regex:"proj_lbl" -s 4 -c 1 -o gpurun_out/prof_labels_${TAG} python bench.py --workload proj_labels --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full_labels_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_labels_${TAG}.log
timeout 600 $FULL -k regex:"proj_ws" -s 4 -c 1 -o gpurun_out/prof_proj_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 2 > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
if [ "$3" != nosan ]; then
  SMALL="tests/test_gpu_labels.py::test_labels_random_vs_oracle[0] tests/test_gpu_labels.py::test_labels_random_vs_oracle[1] tests/test_gpu_labels.py::test_labels_random_vs_oracle[2] tests/test_gpu_parity.py::test_orth_project_random_vs_oracle[0] tests/test_gpu_parity.py::test_orth_project_random_vs_oracle[3] tests/test_gpu_parity.py::test_orth_project_edge_shapes tests/test_gpu_labels.py::test_labels_equal_float_path_and_argument_checks"
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $SMALL -x -q -p no:cacheprovider > gpurun_out/sanitizer_${tool}_${TAG}.log 2>&1
    echo "sanitizer $tool rc=$?"; tail -6 gpurun_out/sanitizer_${tool}_${TAG}.log
  done
fi
