# multi-GPU variants (fence scope of the step's peer stores): usage bash tools/gpu_r2l.sh N TAG
set -x
N=${1:-2}; TAG=${2:-r2l}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
: > $OUT/${TAG}_variants_n$N.txt
port=29700
for fence in sys gpu none; do for rp in 50 0; do
  port=$((port+1))
  TWB200_FENCE=$fence timeout 600 $TR --master-port $port bench.py --gpus $N --steps 2 --warmup 2 --remote-pct $rp 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('fence=$fence remote_pct=$rp n=$N', round(d['value']/1e9,4), 'G ev/s', round(d['ms_per_step'],3), 'ms', 'us/step', round(1e3*d['ms_per_step']/d['supersteps_per_step'],2), hex(int(d['trace_hash_sum'],16)))
" >> $OUT/${TAG}_variants_n$N.txt
done; done
cat $OUT/${TAG}_variants_n$N.txt
TWB200_FENCE=none timeout 900 $TR --master-port 29611 tools/multi_rank_check.py > $OUT/${TAG}_multi_rank_check_fence_none_n$N.txt 2>&1; echo "rc=$?" >> $OUT/${TAG}_multi_rank_check_fence_none_n$N.txt
grep -E "multi-rank|rc=" $OUT/${TAG}_multi_rank_check_fence_none_n$N.txt | tail -8
TWB200_TIME_KERNELS=1 timeout 600 $TR --master-port 29614 bench.py --gpus $N --steps 1 --warmup 1 2> $OUT/${TAG}_time_n$N.err > /dev/null; grep twb200 $OUT/${TAG}_time_n$N.err | tail -2
