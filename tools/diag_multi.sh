j() { python -c "
import sys,json
for t in sys.stdin:
    if t.startswith('{'):
        d=json.loads(t); print('$1', 'G ev/s %.3f'%(d['value']/1e9), 'steps', d['supersteps_per_step'], 'us/step %.1f'%(d['ms_per_step']*1e3/d['supersteps_per_step']), 'remote', d['remote_messages_per_step'])"; }
python bench.py --workload c4 --stop 5 --n-lps 524288 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | j "N=1 512K LPs rp50"
python bench.py --workload c4 --stop 5 --n-lps 131072 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | j "N=1 128K LPs rp50"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
$T 29521 bench.py --gpus 2 --stop 5 --steps 2 --warmup 1 --remote-pct 0 2>/dev/null | j "N=2 1M LPs rp0"
$T 29522 bench.py --gpus 2 --stop 5 --steps 2 --warmup 1 2>/dev/null | j "N=2 1M LPs rp50"
$T 29523 bench.py --gpus 2 --stop 5 --steps 2 --warmup 1 --n-lps 262144 2>/dev/null | j "N=2 256K LPs rp50"
$T 29524 bench.py --gpus 2 --stop 5 --steps 2 --warmup 1 --n-lps 262144 --remote-pct 0 2>/dev/null | j "N=2 256K LPs rp0"
