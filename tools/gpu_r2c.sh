# Round 2, multi-GPU call: staging-ring exchange against the oracle (tools/multi_rank_check.py), then the scaling bench lines.
set -x
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/r2c_topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29611 tools/multi_rank_check.py > $OUT/r2c_multi_rank_check_n$N.txt 2>&1; echo "rc=$?" >> $OUT/r2c_multi_rank_check_n$N.txt
tail -12 $OUT/r2c_multi_rank_check_n$N.txt
for w in c4 c5; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 3 --warmup 3 --workload $w > $OUT/r2c_bench_${w}_n$N.json 2> $OUT/r2c_bench_${w}_n$N.err; tail -c 800 $OUT/r2c_bench_${w}_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 2 --warmup 1 --remote-pct 0 > $OUT/r2c_bench_c4_noremote_n$N.json 2>&1
TWB200_GVT_KERNEL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $N --steps 2 --warmup 1 > $OUT/r2c_bench_c4_gvtkernel_n$N.json 2>&1
for f in $OUT/r2c_bench_c4_n$N.json $OUT/r2c_bench_c5_n$N.json $OUT/r2c_bench_c4_noremote_n$N.json $OUT/r2c_bench_c4_gvtkernel_n$N.json; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$f") if l.startswith("{")][-1])
    print("$f", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "ms", round(d["ms_per_step"],3), d["parity"], d["nvlink"]["achieved_gbs_per_gpu"] if d["nvlink"] else None)
except Exception as e:
    print("$f failed", e)
PY
done
