# sweep of environment-selected launch variants on 1 GPU: TAG then lines "ENVSPEC|bench args" from stdin
OUT=gpurun_out
TAG=${1:-sw}
mkdir -p $OUT
: > $OUT/${TAG}_sweep.txt
while IFS='|' read -r envspec bargs; do
  [ -z "$bargs" ] && continue
  env $envspec timeout 300 python bench.py $bargs --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
ok = False
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); ok = True
        print('$envspec | $bargs |', round(d['value']/1e9,4), 'G ev/s', round(d['ms_per_step'],3), 'ms steps', d['supersteps_per_step'], 'eff', round(d['rollback_efficiency'],4), 'us/step', round(1e3*d['ms_per_step']/d['supersteps_per_step'],2))
if not ok: print('$envspec | $bargs | FAILED')
" >> $OUT/${TAG}_sweep.txt
done
cat $OUT/${TAG}_sweep.txt
