# Round profile recipe (run under gpurun, 1 GPU).  Each ncu pass follows a plain run of the same command.
set -x
OUT=gpurun_out
CMD="python bench.py --workload c4 --stop 2 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c4.csv $CMD > $OUT/ncu_launches_c4.log 2>&1
$CMD > $OUT/plain_c4b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step -s 6 -c 3 -o $OUT/prof_kstep_c4 $CMD > $OUT/ncu_full_c4.log 2>&1
CMD2="python bench.py --workload c3 --stop 10 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > $OUT/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_c3.csv $CMD2 > $OUT/ncu_launches_c3.log 2>&1
