#!/bin/bash
for v in 1 2 3 4; do
  DM_B200_LIB=build/exp/lib_redvar$v.so timeout 120 python scripts/time_proj.py --rows 0 --scene room --steps 100
done
DM_B200_LIB=build/exp/lib_redvar4.so timeout 120 python scripts/time_proj.py --rows 4 --scene room --steps 100
