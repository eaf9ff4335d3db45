"""Profiling helper (not a test): one forward+backward of an EGT block at an arbitrary shape, for `ncu`.

usage: python tools/staged_probe.py N B d d_e h [iters]
"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import egt_b200

DEV = 'cuda:0'


def main():
    N, B, d, de, nh = (int(x) for x in sys.argv[1:6])
    iters = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    torch.manual_seed(0)
    blk = egt_b200.EGTBlock(model_width=d, edge_width=de, num_heads=nh, scale_degree=False, seed=3).to(DEV)
    mask = torch.ones(B, N, dtype=torch.bool, device=DEV)
    h = torch.randn(B, N, d, device=DEV).bfloat16()
    e = torch.randn(B, N, N, de, device=DEV).bfloat16()
    dh, dE = torch.randn_like(h), torch.randn_like(e)
    for _ in range(iters):
        hg, eg = h.clone().requires_grad_(True), e.clone().requires_grad_(True)
        h2, e2 = blk(hg, eg, mask, training=False)
        torch.autograd.backward([h2, e2], [dh, dE])
    torch.cuda.synchronize()
    print('ok', float(h2.float().abs().mean()))


if __name__ == '__main__':
    main()
