# Round 2, call d (2 GPUs): c2/c5 after the tie-sort and cooperative tie resolution, kernel timing of the multi-GPU step.
set -x
OUT=gpurun_out
mkdir -p $OUT
for w in c2 c5 c1 c3 c4; do python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2d_bench_$w.json 2> $OUT/r2d_bench_$w.err; tail -c 300 $OUT/r2d_bench_$w.err; done
for gpw in 3 5 8; do for hz in 0 0.0001; do for b in 4 8; do
  TWB200_GROUPS_PER_WARP=$gpw python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --horizon $hz --batch $b 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('gpw=$gpw hz=$hz batch=$b', round(d['value']/1e9,4), 'G ev/s', round(d['ms_per_step'],1), 'ms steps', d['supersteps_per_step'], 'eff', round(d['rollback_efficiency'],3))
" >> $OUT/r2d_c2_sweep.txt
done; done; done
cat $OUT/r2d_c2_sweep.txt
python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-group-kernel --horizon 2 > $OUT/r2d_bench_c2_kstep.json 2>&1
TWB200_TIME_KERNELS=1 python bench.py --n-lps 524288 --steps 1 --warmup 1 --no-cpu-baseline 2> $OUT/r2d_time_1gpu_512k.err > /dev/null; grep twb200 $OUT/r2d_time_1gpu_512k.err
TWB200_TIME_KERNELS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 1 --warmup 1 2> $OUT/r2d_time_n2.err > $OUT/r2d_time_n2.json; grep twb200 $OUT/r2d_time_n2.err
TWB200_TIME_KERNELS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 1 --warmup 1 --remote-pct 0 2> $OUT/r2d_time_n2_noremote.err > /dev/null; grep twb200 $OUT/r2d_time_n2_noremote.err
for w in c2 c5 c1 c3 c4 c2_kstep; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/r2d_bench_$w.json") if l.startswith("{")][-1])
    print("$w", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"], 4), d["parity"].get("checked"))
except Exception as e:
    print("$w failed", e)
PY
done
