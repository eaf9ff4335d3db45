"""Cycles per small tcgen05.mma (M=128, K=16) as issued by the fused kernels: operand source, N, chain structure.
usage (GPU box): python tools/mma_timing.py"""
import ctypes as C
import sys

sys.path.insert(0, '.')
import torch                                              # noqa: E402  (CUDA context)
from egt_b200 import _lib as L                            # noqa: E402

lib = L.load()
torch.zeros(1, device='cuda')
out = (C.c_longlong * 2)()
names = {0: 'A smem K-major', 1: 'A tensor memory', 2: 'A smem MN-major'}
print('mode               N  ksteps chains ndst   cycles/mma (issue->done)   issue cycles/mma')
for a_mode in (0, 1, 2):
    for N in (16, 32, 64, 128):
        for ksteps, chains, ndst in ((8, 64, 1), (8, 64, 2), (1, 512, 1)):
            rc = lib.egt_debug_mma_timing(a_mode, N, ksteps, chains, ndst, out, None)
            assert rc == 0, L.last_error() if hasattr(L, 'last_error') else rc
            n = ksteps * chains
            print(f'{names[a_mode]:18s} {N:3d} {ksteps:6d} {chains:6d} {ndst:4d}   {out[0] / n:8.1f}                   {out[1] / n:8.1f}')
