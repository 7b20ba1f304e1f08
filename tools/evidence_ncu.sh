#!/bin/bash
# ncu --set full captures of the HBM-bound and exact-tier kernels, run on
# the GPU box:   bash tools/evidence_ncu.sh    -> gpurun_out/ncu_<kernel>_r02.{ncu-rep,txt}
set -x
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor_op_dmma.sum"
cap () {  # name, kernel regex, skip, command...
  name=$1; shift; rx=$1; shift; skip=$1; shift
  $NCU -k regex:$rx -s $skip -c 1 -o $OUT/ncu_${name}_r02 -f "$@" > $OUT/ncu_${name}_r02.log 2>&1
  ncu -i $OUT/ncu_${name}_r02.ncu-rep --page raw --csv > $OUT/ncu_${name}_r02_raw.csv 2>/dev/null
}
python tools/bench_rotation.py > $OUT/bench_rotation_r02.log 2>&1
cap rotate_exact k_rotate_assemble 2 python tools/bench_rotation.py
cap rotate_between k_rotate_assemble 8 python tools/bench_rotation.py
cap lerp_rows k_lerp_rows 2 python tools/bench_rotation.py
cap single_fascicle k_single_fascicle 1 python tools/bench_kernels_small.py single
cap pairs3 'k_pairs' 1 python tools/bench_kernels_small.py pairs3
