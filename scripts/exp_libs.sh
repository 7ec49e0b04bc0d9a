#!/bin/bash
# Experiment helper: one bench line per library variant (build/exp/lib_<name>.so), printed as "name ms/step kernel-ms value".
# usage: scripts/exp_libs.sh <workload> <steps> name...
WL=$1; STEPS=$2; shift 2
for l in "$@"; do
  DM_B200_LIB=build/exp/lib_$l.so timeout 300 python bench.py --workload $WL --steps $STEPS --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$l', round(d['ms_per_step'],4), round(r.get('kernel_ms_per_step', d['ms_per_step']),4), round(d['value']))"
done
