run() {  # env workload-args
  env $1 python bench.py --workload $2 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
t=sys.stdin.read().strip().split('\n')[-1]
try:
    d=json.loads(t); print('$1','$2','G ev/s %.3f'%(d['value']/1e9),'e2e %.3f'%(d['e2e']['value']/1e9),'steps',d['supersteps_per_step'],'eff %.4f'%d['rollback_efficiency'],'us/step %.1f'%(d['ms_per_step']*1e3/d['supersteps_per_step']))
except Exception as e: print('$1 $2 FAILED', t[:200])"
}
for t in 1 2 4 8; do run TWB200_TILES=$t "c4 --stop 5"; done
for t in 1 2; do run TWB200_TILES=$t "c3 --stop 30"; run TWB200_TILES=$t "c5 --stop 5"; done
run TWB200_TILES=4 "c4 --stop 5 --horizon 0.3 --batch 8"
run TWB200_TILES=4 "c4 --stop 5 --horizon 0.4 --batch 6"
run TWB200_TILES=4 "c4 --stop 5 --horizon 0.2 --batch 6"
run TWB200_TILES=1 "c2 --stop 200"
