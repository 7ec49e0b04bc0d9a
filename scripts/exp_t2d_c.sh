#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_layouts or orth_project" 2>&1 | tail -3
for rows in ${ROWS:-0 4}; do
  for scene in room iid; do
    timeout 120 python scripts/time_proj.py --rows $rows --scene $scene --steps 100
  done
  timeout 120 python scripts/time_proj.py --rows $rows --scene room --steps 20 --hw 720x1280 --c 40 --b 32
  timeout 120 python scripts/time_proj.py --rows $rows --scene room --steps 100 --c 0
done
