# Round 2, second GPU call (1 GPU): closed-group kernel operating points (groups per warp x window x batch), ncu captures.
set -x
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/r2b_c2_sweep.txt
for gpw in 2 5 10; do for hz in 0 0.0001 0.05 2; do for b in 1 2 4; do
  TWB200_GROUPS_PER_WARP=$gpw python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --horizon $hz --batch $b 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('gpw=$gpw hz=$hz batch=$b', round(d['value']/1e9,4), 'G ev/s', round(d['ms_per_step'],1), 'ms steps', d['supersteps_per_step'], 'eff', round(d['rollback_efficiency'],3))
" >> $OUT/r2b_c2_sweep.txt
done; done; done
cat $OUT/r2b_c2_sweep.txt
# ncu: k_run_groups (c2, short), k_step c5 (short), k_step c4 (short) -- each after a plain run of the same command
CMD="python bench.py --workload c2 --stop 50 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD > $OUT/r2b_plain_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_run_groups -c 1 -f -o $OUT/r2b_prof_groups_c2 $CMD > $OUT/r2b_ncu_c2.log 2>&1
CMD="python bench.py --workload c5 --stop 1 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD > $OUT/r2b_plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_step -s 10 -c 1 -f -o $OUT/r2b_prof_kstep_c5 $CMD > $OUT/r2b_ncu_c5.log 2>&1
CMD="python bench.py --workload c4 --stop 2 --steps 1 --warmup 0 --no-cpu-baseline"
$CMD > $OUT/r2b_plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 1 -f -o $OUT/r2b_prof_kstep_c4 $CMD > $OUT/r2b_ncu_c4.log 2>&1
ls -la $OUT/*.ncu-rep
