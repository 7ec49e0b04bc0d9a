#!/bin/bash
# ablations of the tile layouts: load pipeline + phase A + resolve only (nob), without REDs (nored), L2 promotion
for lib in nob nored promo0 promo1; do
  for rows in 0 4; do
    DM_B200_LIB=build/exp/lib_$lib.so timeout 120 python scripts/time_proj.py --rows $rows --scene room --steps 100
  done
done
DM_B200_LIB=build/exp/lib_nob.so timeout 120 python scripts/time_proj.py --rows 8 --scene room --steps 100
DM_B200_LIB=build/exp/lib_nored.so timeout 120 python scripts/time_proj.py --rows 8 --scene room --steps 100
