"""Peer-memory gradient all-reduce (csrc/peer_allreduce.cu) against NCCL on >= 2 GPUs of one node; eager
calls and CUDA-graph replay.  Skipped on a single-GPU box (the host-side sharding logic is covered on
CPU/gloo by tests/test_dp_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
@pytest.mark.parametrize('protocol', ['push', 'pull'])
def test_peer_allreduce_matches_nccl(protocol):
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr',
           '127.0.0.1', '--master-port', '29541', os.path.join(ROOT, 'tools', 'peer_allreduce_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=dict(os.environ, EGT_PEER_PROTOCOL=protocol))
    assert r.returncode == 0 and 'PEER_ALLREDUCE_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
