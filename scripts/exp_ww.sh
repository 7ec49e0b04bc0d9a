#!/bin/bash
# Experiment: consumer warps per CTA of the projection kernel (DM_WS_WARPS override), config 2 and config 5 shapes.
for ww in ${WWS:-8 6 4 3 2}; do
  for scene in room iid; do
    DM_WS_WARPS=$ww timeout 200 python scripts/time_proj.py --scene $scene --steps 100 2>&1 | tail -1 | sed "s/^/ww=$ww /"
  done
  DM_WS_WARPS=$ww timeout 300 python scripts/time_proj.py --scene room --steps 20 --hw 720x1280 --c 40 --b 32 2>&1 | tail -1 | sed "s/^/ww=$ww cfg5 /"
done
