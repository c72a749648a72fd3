# scaling evidence at N GPUs: bench lines of c4 (headline) and c5, parity against the oracle asserted by bench.py itself
set -x
N=${1:-2}; TAG=${2:-r2y}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_c4_n$N.json 2> $OUT/${TAG}_bench_c4_n$N.err; tail -c 400 $OUT/${TAG}_bench_c4_n$N.err
timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 2 --warmup 2 --workload c5 > $OUT/${TAG}_bench_c5_n$N.json 2> $OUT/${TAG}_bench_c5_n$N.err; tail -c 400 $OUT/${TAG}_bench_c5_n$N.err
if [ "$3" = "check" ]; then timeout 900 $TR --master-port 29611 tools/multi_rank_check.py > $OUT/${TAG}_multi_rank_check_n$N.txt 2>&1; echo "rc=$?" >> $OUT/${TAG}_multi_rank_check_n$N.txt; grep -E "multi-rank|rc=" $OUT/${TAG}_multi_rank_check_n$N.txt | tail -7; fi
for w in c4 c5; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_${w}_n$N.json") if l.startswith("{")][-1])
    print("$w n=$N", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "e2e(dev pop)", round(d["e2e_device_population"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "ms", round(d["ms_per_step"], 3), "us/step", round(1e3*d["ms_per_step"]/d["supersteps_per_step"], 2), [k for k, v in d["parity"].items() if v is False], "nvlink GB/s/gpu", d["nvlink"]["achieved_gbs_per_gpu"] if d["nvlink"] else None)
except Exception as e:
    print("$w failed", e)
PY
done
