#!/bin/bash
# usage: tools/ncu_variant.sh lib.so tag   -> gpurun_out/ncu_<tag>.csv (sn_kernel, one launch, selected metrics)
REDBACK_B200_LIB=$(realpath $1) ncu --clock-control none -k regex:sn_kernel -c 6 --csv \
  --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum \
  python bench.py --samples 400000 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -v "^==" > gpurun_out/ncu_$2.csv
