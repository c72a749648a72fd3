# usage: bash tools/sweep_window.sh  -- optimism window x batch sweep of the default library, then launch-shape variants
# (aivika-distributed_b200/csrc/variants/lib_b<block>_m<minblocks>_c<candidates>.so) at one operating point.
run() {  # lib workload-args hz batch
  TWB200_LIB=$1 python bench.py --workload $2 --horizon $3 --batch $4 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
t=sys.stdin.read().strip().split('\n')[-1]
try:
    d=json.loads(t); print('$1'.split('/')[-1],'$2','hz',$3,'batch',$4,'G ev/s %.3f'%(d['value']/1e9),'e2e %.3f'%(d['e2e']['value']/1e9),'steps',d['supersteps_per_step'],'eff %.4f'%d['rollback_efficiency'],'us/step %.1f'%(d['ms_per_step']*1e3/d['supersteps_per_step']))
except Exception as e: print('$1 $2 hz $3 batch $4 FAILED', t[:200])"
}
LIB=$PWD/aivika-distributed_b200/csrc/libtwb200.so
for hz in 0.2 0.3 0.4 0.5; do for b in 4 6 8 16; do run $LIB "c4 --stop 5" $hz $b; done; done
for hz in 0.2 0.3; do for b in 4 6 8; do run $LIB "c3 --stop 30" $hz $b; done; done
for hz in 0.3 0.5; do for b in 8 16 32; do run $LIB "c5 --stop 5" $hz $b; done; done
for v in aivika-distributed_b200/csrc/variants/lib_*.so; do run $PWD/$v "c4 --stop 5" 0.3 8; run $PWD/$v "c3 --stop 30" 0.2 6; done
