# ncu capture of k_step on the skewed workload c5 (hot LPs: whole-region search, slow drain)
set -x
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --workload c5 --stop 6 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD > $OUT/c5prof_plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_step -s 70 -c 1 -f -o $OUT/c5prof_prof_kstep_c5 $CMD > $OUT/c5prof_ncu_c5.log 2>&1
tail -2 $OUT/c5prof_plain_c5.log | cut -c1-400
ls -la $OUT/c5prof_prof_kstep_c5.ncu-rep
