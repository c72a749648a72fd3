for f in aivika-distributed_b200/csrc/variants/*.so; do
  for w in "c3 --stop 30" "c4 --stop 5"; do
    TWB200_LIB=$PWD/$f python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$f','$w', 'G ev/s %.3f'%(d['value']/1e9), 'e2e %.3f'%(d['e2e']['value']/1e9), 'steps',d['supersteps_per_step'],'eff %.4f'%d['rollback_efficiency'])"
  done
done
