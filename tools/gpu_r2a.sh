# Round 2, first GPU call (1 GPU): parity tests, bench lines of every configuration, 128K-shard occupancy variants.
set -x
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/r2a_pytest_gpu.log
tail -5 $OUT/r2a_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > $OUT/r2a_bench_c4.json 2> $OUT/r2a_bench_c4.err; tail -c 600 $OUT/r2a_bench_c4.err
for w in c2 c3 c5 c1; do python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2a_bench_$w.json 2> $OUT/r2a_bench_$w.err; tail -c 400 $OUT/r2a_bench_$w.err; done
python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-group-kernel > $OUT/r2a_bench_c2_kstep.json 2>&1
# the 8-GPU share of c4 on one GPU (128K LPs): does one wave (more resident warps) pay?
for v in "" _mb5 _mb6 _mb8; do
  TWB200_LIB=$PWD/aivika-distributed_b200/csrc/libtwb200$v.so python bench.py --n-lps 131072 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('variant$v n_lps=131072', d['value'], d['ms_per_step'], d['supersteps_per_step'])
" >> $OUT/r2a_variants.txt
  TWB200_LIB=$PWD/aivika-distributed_b200/csrc/libtwb200$v.so python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('variant$v c4', d['value'], d['ms_per_step'], d['supersteps_per_step'])
" >> $OUT/r2a_variants.txt
done
cat $OUT/r2a_variants.txt
for w in c4 c2 c3 c5 c1; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/r2a_bench_$w.json") if l.startswith("{")][-1])
    print("$w", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "frac", round(d["roofline"]["frac"], 4), d["parity"])
except Exception as e:
    print("$w failed", e)
PY
done
