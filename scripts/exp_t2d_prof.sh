#!/bin/bash
for lib in prof profnored; do
  for rows in 0 4; do
    DM_PROFILE=1 DM_PROJ_KERNEL=w DM_B200_LIB=build/exp/lib_$lib.so timeout 120 python scripts/time_proj.py --rows $rows --scene room --steps 50
  done
done
