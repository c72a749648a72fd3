# Round profile recipe (run under gpurun, 1 GPU).  Each ncu pass follows a plain run of the same command.
# Outputs: launch list of c4 (gpu__time_duration per launch) and one `ncu --set full` capture of the dominant kernel of
# every benched configuration (c4, c3, c5, c2) and of a 128K-LP partition (one GPU's share of c4 at 8 GPUs); read them
# afterwards with tools/ncu_traffic.py (-> profiles/traffic.json) and tools/ncu_lines.py.
set -x
OUT=gpurun_out
TAG=${1:-r2}
mkdir -p $OUT
cap() {   # name, kernel regex, launches to skip, bench args...
  local name=$1 rx=$2 skip=$3; shift 3
  local CMD="python bench.py $* --steps 1 --warmup 0 --no-cpu-baseline"
  $CMD > $OUT/${TAG}_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/${TAG}_prof_$name $CMD > $OUT/${TAG}_ncu_$name.log 2>&1
}
CMD="python bench.py --workload c4 --stop 2 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain_c4l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_c4.csv $CMD > $OUT/${TAG}_ncu_launches_c4.log 2>&1
cap kstep_c4 k_step 12 --workload c4 --stop 2
cap kstep_c3 k_step 40 --workload c3 --stop 10
cap kstep_c5 k_step 70 --workload c5 --stop 6
cap groups_c2 k_run_groups 0 --workload c2 --stop 50
cap kstep_c4_128k k_step 12 --n-lps 131072 --stop 2
ls -la $OUT/${TAG}_prof_*.ncu-rep
