#!/bin/bash
# Evidence run on the GPU box:  bash tools/evidence_ncu.sh
#   1. compute-sanitizer memcheck / racecheck / synccheck on tools/sanitize_small.py
#   2. ncu launch list of a short bench.py run (kernel time shares)
#   3. ncu --set full captures: k_fast_tiles<0,0> / <1,0>, k_rotate_assemble (exact-G and
#      between-shell), k_lerp_rows, k_single_fascicle, k_pairs<3>
# Everything lands in gpurun_out/ (scratch); tools/ncu_summary.py condenses the raw pages for profiles/.
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $OUT/sanitizer_${tool}_r02.log 2>&1
  echo "$tool rc=$?" >> $OUT/sanitizer_${tool}_r02.log
  tail -4 $OUT/sanitizer_${tool}_r02.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_r02.csv \
    python bench.py --voxels 20000 --steps 2 --warmup 3 --no-extra > $OUT/launches_r02.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
cap () {  # name, kernel regex, skip, count, command...
  name=$1; shift; rx=$1; shift; skip=$1; shift; cnt=$1; shift
  timeout 900 $NCU -k regex:$rx -s $skip -c $cnt -o $OUT/ncu_${name}_r02 -f "$@" > $OUT/ncu_${name}_r02.log 2>&1
  ncu -i $OUT/ncu_${name}_r02.ncu-rep --page raw --csv > $OUT/ncu_${name}_r02_raw.csv 2>/dev/null
  ls -la $OUT/ncu_${name}_r02.ncu-rep
}
cap fast_tiles k_fast_tiles 6 2 python bench.py --voxels 6000 --steps 1 --warmup 3 --no-extra
python tools/bench_rotation.py > $OUT/bench_rotation_r02.log 2>&1; cat $OUT/bench_rotation_r02.log
cap rotate_exact k_rotate_assemble 2 1 python tools/bench_rotation.py
cap rotate_between k_rotate_assemble 8 1 python tools/bench_rotation.py
cap lerp_rows k_lerp_rows 2 1 python tools/bench_rotation.py
cap single_fascicle k_single_fascicle 1 1 python tools/bench_kernels_small.py single
cap pairs3 'k_pairs' 1 1 python tools/bench_kernels_small.py pairs3
