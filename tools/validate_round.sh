# Round-end validation on one GPU: parity suite, smoke, bench line of every configuration, the reference arm.
set -x
TAG=${1:-r2x}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err; tail -c 400 $OUT/${TAG}_bench_c4.err
for w in c3 c5 c2 c1; do timeout 600 python bench.py --workload $w > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; tail -c 300 $OUT/${TAG}_bench_$w.err; done
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_ref_c4.json 2>&1
for w in c4 c3 c5 c2 c1; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_$w.json") if l.startswith("{")][-1])
    print("$w", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "e2e(dev pop)", round(d["e2e_device_population"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], [k for k, v in d["parity"].items() if v is False], d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("$w failed", e)
PY
done
tail -c 600 $OUT/${TAG}_bench_ref_c4.json
