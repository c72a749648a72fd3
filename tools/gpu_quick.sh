# quick 1-GPU bench lines (no tests, no ncu): TAG [workloads...]
set -x
OUT=gpurun_out
TAG=${1:-q}; shift
mkdir -p $OUT
for w in ${@:-c4 c3}; do timeout 600 python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; tail -c 300 $OUT/${TAG}_bench_$w.err; done
timeout 300 python bench.py --n-lps 131072 --steps 2 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_bench_c4_128k.json 2>&1
for w in ${@:-c4 c3} c4_128k; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_$w.json") if l.startswith("{")][-1])
    print("$w", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["parity"].get("checked"), [k for k,v in d["parity"].items() if v is False])
except Exception as e:
    print("$w failed", e)
PY
done
