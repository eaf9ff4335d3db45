"""One forward + backward of the feed-forward half at a headline shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import egt_b200

kind = sys.argv[1] if len(sys.argv) > 1 else 'edge8'
name, shape, w = {'edge8': ('edge', (128, 128, 128, 8), 8), 'node64': ('node', (128, 128, 64), 64),
                  'edge32': ('edge', (16, 256, 256, 32), 32), 'edge64': ('edge', (128, 64, 64, 64), 64)}[kind]
dev = 'cuda:0'
ffn = egt_b200.EGTFFN(w, channel=name).to(dev)
x = torch.randn(*shape, device=dev).bfloat16().requires_grad_(True)
dy = torch.randn(*shape, device=dev).bfloat16()
for _ in range(3):
    y = ffn(x); torch.autograd.grad(y, [x, ffn.flat], dy)
torch.cuda.synchronize()
