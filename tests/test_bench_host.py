"""Host-side pieces of bench.py that need no GPU: the algorithmic byte / FLOP formulas of SURVEY.md section 8(d)
against its worked numbers, and the JSON line of the CPU reference arm on a tiny sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_and_flops_match_the_survey():
    # SURVEY.md 8(d) "Worked numbers" (bf16): C0 1.43 MB / 44 MFLOP, C2 0.52 MB, C4 3.00 MB, C5 84.8 MB / 3.02 GFLOP
    mb = lambda w: bench.alg_bytes_per_graph(*(bench.WORKLOADS[w][k] for k in ('N', 'd', 'd_e', 'h'))) / 1e6
    gf = lambda w: bench.alg_flops_per_graph(*(bench.WORKLOADS[w][k] for k in ('N', 'd', 'd_e', 'h'))) / 1e9
    assert abs(mb('C0') - 1.43) < 0.01 and abs(mb('C2') - 0.52) < 0.01
    assert abs(mb('C4') - 3.00) < 0.01 and abs(mb('C5') - 84.8) < 0.1
    assert abs(gf('C0') - 0.044) < 0.001 and abs(gf('C5') - 3.02) < 0.01
    for w in bench.WORKLOADS.values():
        a = [w[k] for k in ('N', 'd', 'd_e', 'h')]
        assert bench.alg_bytes_fwd(*a) + bench.alg_bytes_bwd(*a) == bench.alg_bytes_per_graph(*a)


def test_reference_arm_prints_the_contract_line():
    r = bench.cpu_reference_run(dict(N=16, d=64, d_e=8, h=8, B=4), steps=2, warmup=1, sample_graphs=2, budget_s=5.0)
    assert r['value'] > 0 and r['kind'] == 'port' and r['cores'] >= 1 and r['steps'] == 2
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'C1',
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, RANK='0'))
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == bench.METRIC and line['unit'] == 'graphs/s'
    assert line['gpu_launches'] == 0 and line['e2e']['h2d_bytes_per_step'] == 0
    assert line['cpu_baseline']['kind'] == 'port' and line['config']['workload'] == 'C1'
    # ranks other than 0 stay silent under torchrun
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'C1',
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, RANK='1'))
    assert out.returncode == 0 and out.stdout.strip() == ''
