#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_layouts" 2>&1 | tail -3
for rows in 0 4 8; do
  for scene in room iid; do
    timeout 120 python scripts/time_proj.py --rows $rows --scene $scene --steps 100
  done
done
timeout 120 python scripts/time_proj.py --rows 4 --scene room --steps 20 --hw 720x1280 --c 40 --b 32
for rows in 4; do
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -f -k regex:proj_ws -s 5 -c 1 -o gpurun_out/src_rows${rows}_${TAG:-x} python scripts/time_proj.py --rows $rows --scene room --steps 3 2>&1 | tail -2
done
