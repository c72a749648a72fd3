for mb in 4 6 8; do
 for w in "c3 --stop 30" "c4 --stop 5"; do
  for hz in 0.1 0.2 0.4; do
    TWB200_LIB=$PWD/aivika-distributed_b200/csrc/variants/lib_mb$mb.so python bench.py --workload $w --horizon $hz --batch $mb --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$w','hz',$hz,'batch',$mb, 'G ev/s %.3f'%(d['value']/1e9), 'steps',d['supersteps_per_step'],'eff %.4f'%d['rollback_efficiency'])"
  done
 done
done
