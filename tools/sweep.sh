#!/bin/bash
# Per-workload sweep on one B200: every BASELINE.json configuration, the N sweep at the headline widths
# (S64..S512: d=64, d_e=8, h=8) and at the widths of config 5 (W64..C5: d=128, d_e=32, h=16).
#   gpurun --timeout 1500 -- 'bash tools/sweep.sh'
# Writes the raw bench.py lines to gpurun_out/sweep.jsonl (copy to profiles/rN_sweep.jsonl) and prints one summary
# line per workload: path, graphs/s, ms/step, roofline fraction of the whole step, end-to-end graphs/s, per-kernel
# microseconds (CUDA events, eager pass).
out=gpurun_out/sweep.jsonl
mkdir -p gpurun_out
: > $out
for w in C0 C0e64 C1 C1s C2 C3 C4 C5 S64 S256 S512 W64 W128 W256; do
  timeout 240 python bench.py --workload $w --steps ${SWEEP_STEPS:-10} --warmup 3 --no-cpu --input-sets 2 2>gpurun_out/sweep_$w.err | tail -1 > gpurun_out/sweep_$w.json
  cat gpurun_out/sweep_$w.json >> $out
  python - "$w" <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/sweep_{sys.argv[1]}.json').read())
    k = d['roofline']['kernels']
    t = d.get('training_step') or {}
    fl = d.get('full_layer') or {}
    print(d['config']['workload'], d['config'].get('path'), round(d['value']), 'graphs/s', round(d['ms_per_step'], 4), 'ms',
          'step_frac', round(d['roofline']['step_frac'], 3), 'e2e', round(d['e2e']['value']),
          'train', round(t.get('value', 0)), 'layer_ms', round(fl.get('ms_per_step', 0), 4),
          ' '.join(f"{n.replace('_kernel', '')}={v['ms_total'] / v['launches'] * 1000:.0f}" for n, v in k.items()))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
PY
done
