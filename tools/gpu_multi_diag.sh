# Round 2, multi-GPU call: staging-ring exchange against the oracle (tools/multi_rank_check.py), scaling bench line of c4,
# the same without remote messages, per-kernel timing.  usage: gpurun --gpus N -- bash tools/gpu_multi_diag.sh N TAG
set -x
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29611 tools/multi_rank_check.py > $OUT/${TAG}_multi_rank_check_n$N.txt 2>&1; echo "rc=$?" >> $OUT/${TAG}_multi_rank_check_n$N.txt
grep -E "multi-rank|rc=" $OUT/${TAG}_multi_rank_check_n$N.txt | tail -8
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_c4_n$N.json 2> $OUT/${TAG}_bench_c4_n$N.err; tail -c 600 $OUT/${TAG}_bench_c4_n$N.err
timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 2 --warmup 2 --remote-pct 0 > $OUT/${TAG}_bench_c4_noremote_n$N.json 2>&1
TWB200_TIME_KERNELS=1 timeout 600 $TR --master-port 29614 bench.py --gpus $N --steps 1 --warmup 1 2> $OUT/${TAG}_time_n$N.err > $OUT/${TAG}_time_n$N.json; grep twb200 $OUT/${TAG}_time_n$N.err | tail -$N
TWB200_TIME_KERNELS=1 timeout 600 $TR --master-port 29615 bench.py --gpus $N --steps 1 --warmup 1 --remote-pct 0 2> $OUT/${TAG}_time_noremote_n$N.err > /dev/null; grep twb200 $OUT/${TAG}_time_noremote_n$N.err | tail -$N
for f in $OUT/${TAG}_bench_c4_n$N.json $OUT/${TAG}_bench_c4_noremote_n$N.json; do python - <<PY
import json
try:
    d = json.loads([l for l in open("$f") if l.startswith("{")][-1])
    print("$f", round(d["value"]/1e9, 4), "G ev/s  e2e", round(d["e2e"]["value"]/1e9, 4), "eff", round(d["rollback_efficiency"], 4), "steps", d["supersteps_per_step"], "ms", round(d["ms_per_step"],3), "us/step", round(1e3*d["ms_per_step"]/d["supersteps_per_step"],2), d["parity"], d["nvlink"]["achieved_gbs_per_gpu"] if d["nvlink"] else None)
except Exception as e:
    print("$f failed", e)
PY
done
