timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in C1 C3 C5; do timeout 120 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --input-sets 2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print(d['config']['workload'], d['config'].get('path'), round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), ' '.join(f\"{n.replace('_kernel','')}={v['ms_total']/v['launches']*1000:.0f}\" for n,v in k.items()))
"; done
