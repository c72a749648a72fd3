# Round profile recipe (run under gpurun, 1 GPU).  Each ncu pass follows a plain run of the same command.
# c4 --stop 2 takes 16 super-steps; tw_run keeps two graphs of 16 k_step launches in flight, so a run issues 32 launches
# of which the last 16 return at once (GVT past the stop time): launches 32..47 are the real steps of the second run.
set -x
OUT=gpurun_out
TAG=${1:-r1b}
CMD="python bench.py --workload c4 --stop 2 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_c4.csv $CMD > $OUT/${TAG}_ncu_launches_c4.log 2>&1
$CMD > $OUT/${TAG}_plain_c4b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step -s 36 -c 2 -f -o $OUT/${TAG}_prof_kstep_c4 $CMD > $OUT/${TAG}_ncu_full_c4.log 2>&1
CMD2="python bench.py --workload c3 --stop 10 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > $OUT/${TAG}_plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_c3.csv $CMD2 > $OUT/${TAG}_ncu_launches_c3.log 2>&1
