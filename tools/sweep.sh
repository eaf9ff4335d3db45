#!/bin/bash
# Reproduces the per-workload table of README.md / DESIGN.md on one B200:
#   gpurun --timeout 900 -- 'bash tools/sweep.sh'
# One line per workload: path, graphs/s, ms/step, end-to-end graphs/s, per-kernel microseconds (CUDA events, eager pass).
for w in C0 C1 C2 C3 C4 C5 S64 S256 S512; do
  timeout 180 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --input-sets 2 2>&1 | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read()); k = d['roofline']['kernels']
print(d['config']['workload'], d['config'].get('path'), round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']),
      ' '.join(f\"{n.replace('_kernel', '')}={v['ms_total'] / v['launches'] * 1000:.0f}\" for n, v in k.items()))
"
done
