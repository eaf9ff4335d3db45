"""The small weight-only / gradient-fold launches of a call run on a side stream by default (csrc/abi.cu: SideStream).
EGT_SIDE_STREAM=0 keeps every launch on the caller's stream; the switch is read once per process, so the single-stream
variant is exercised in a child process on shapes that use every forked branch (cuBLAS node side, tcgen05 node side,
narrow and wide fused kernels)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_stream_variant_passes_the_full_size_parity_cases():
    env = dict(os.environ, EGT_SIDE_STREAM='0')
    cmd = [sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_parity_gpu.py'), '-q', '-x', '-k',
           'full_size_vs_oracle_sample and (190-96 or 128-64 or 37-64 or 512-128)']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and ' passed' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
