B="python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e"
for st in S1 S2; do
    $B --settings $st 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$st', d['config'].get('kernel'), 'ms', round(d['ms_per_step'],3), 'iters', d.get('admm_iters_per_step'))
"
done
$B --workload config2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('config2', d['config'].get('kernel'), 'ms', round(d['ms_per_step'],4))
"
$B --fp32 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fp32', d['config'].get('kernel'), 'ms', round(d['ms_per_step'],4))
"
