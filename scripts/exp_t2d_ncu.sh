#!/bin/bash
M=gpu__time_duration.sum,smsp__inst_executed_op_global_red.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum,lts__t_sector_op_red_hit_rate.pct,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum
for rows in ${ROWS:-0 4}; do
  ncu --metrics $M --clock-control none -k regex:proj_ws -s 5 -c 1 python scripts/time_proj.py --rows $rows --scene ${SCENE:-room} --steps 3 2>&1 | grep -E "proj_ws_kernel|gpu__time|smsp__|lts__|l1tex__|sm__" | cut -c1-150
done
