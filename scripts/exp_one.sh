#!/bin/bash
# usage: LIBS="a b" ROWS="0" bash scripts/exp_one.sh
for lib in $LIBS; do
  for rows in ${ROWS:-0}; do
    for scene in room iid; do
      DM_B200_LIB=build/exp/lib_$lib.so timeout 120 python scripts/time_proj.py --rows $rows --scene $scene --steps 100
    done
  done
done
