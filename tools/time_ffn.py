"""Timing helper: feed-forward half of a layer at the headline shapes (CUDA events, warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import egt_b200
from egt_b200 import _lib as L

dev = 'cuda:0'
for name, shape, w in (('edge', (128, 128, 128, 8), 8), ('node', (128, 128, 64), 64), ('edge', (16, 256, 256, 32), 32),
                       ('edge', (128, 64, 64, 64), 64), ('node', (32, 512, 128), 128)):
    ffn = egt_b200.EGTFFN(w, channel=name).to(dev)
    x = torch.randn(*shape, device=dev).bfloat16().requires_grad_(True)
    dy = torch.randn(*shape, device=dev).bfloat16()
    for _ in range(3):
        y = ffn(x); torch.autograd.grad(y, [x, ffn.flat], dy)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(10):
        y = ffn(x)
    e[1].record()
    for _ in range(10):
        y = ffn(x); torch.autograd.grad(y, [x, ffn.flat], dy)
    e[2].record()
    torch.cuda.synchronize()
    f = e[0].elapsed_time(e[1]) / 10
    fb = e[1].elapsed_time(e[2]) / 10
    print(f'{name} FFN {shape}: fwd {f*1e3:.1f} us, fwd+bwd {fb*1e3:.1f} us', flush=True)
    lib = L.load()
    lib.egt_profile_enable(1)
    for _ in range(5):
        y = ffn(x); torch.autograd.grad(y, [x, ffn.flat], dy)
    torch.cuda.synchronize()
    prof = L.profile_read()
    lib.egt_profile_enable(0)
    print('   kernels: ' + ', '.join(f'{k} {v[0] / v[1] * 1e3:.1f} us' for k, v in prof.items()), flush=True)
