# Round 2, index-ring kernel (1 GPU): parity tests, bench lines, launch list + full ncu capture of k_step on c4.
set -x
OUT=gpurun_out
TAG=${1:-r2f}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err; tail -c 600 $OUT/${TAG}_bench_c4.err
for w in c3 c2 c5; do timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; tail -c 400 $OUT/${TAG}_bench_$w.err; done
timeout 300 python bench.py --n-lps 131072 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_c4_128k.json 2>&1
for w in c4 c3 c2 c5 c4_128k; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_$w.json") if l.startswith("{")][-1])
    print("$w", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["parity"])
except Exception as e:
    print("$w failed", e)
PY
done
CMD="python bench.py --workload c4 --stop 2 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_c4.csv $CMD > $OUT/${TAG}_ncu_launches_c4.log 2>&1
$CMD > $OUT/${TAG}_plain_c4b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_step -s 36 -c 1 -f -o $OUT/${TAG}_prof_kstep_c4 $CMD > $OUT/${TAG}_ncu_full_c4.log 2>&1
ls -la $OUT/*.ncu-rep
