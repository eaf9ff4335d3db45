"""Bring-up helper (not a test): fused tcgen05 path vs staged kernels on the GPU, per-output error report."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import egt_b200
from egt_b200 import _lib as L

DEV = 'cuda:0'


def run(N, B, training=False, rmp=0.0, scale_degree=True, seed=0):
    lib = L.load()
    d, de, nh = 64, 8, 8
    torch.manual_seed(seed)
    blk = egt_b200.EGTBlock(model_width=d, edge_width=de, num_heads=nh, scale_degree=scale_degree,
                            random_mask_prob=rmp, seed=3).to(DEV)
    with torch.no_grad():
        blk.flat.add_(0.05 * torch.randn_like(blk.flat))
    nn_ = torch.randint(max(1, N // 2), N + 1, (B,))
    mask = (torch.arange(N)[None] < nn_[:, None]).to(DEV)
    h = torch.randn(B, N, d).bfloat16().to(DEV)
    e = torch.randn(B, N, N, de).bfloat16().to(DEV)
    dh = torch.randn(B, N, d).bfloat16().to(DEV)
    dE = torch.randn(B, N, N, de).bfloat16().to(DEV)
    outs = {}
    for force in (1, 0):
        lib.egt_debug_force_staged(force)
        blk.rng.offset = 10
        hg, eg = h.clone().requires_grad_(True), e.clone().requires_grad_(True)
        blk.flat.grad = None
        h2, e2 = blk(hg, eg, mask, training=training)
        torch.autograd.backward([h2, e2], [dh, dE])
        torch.cuda.synchronize()
        outs[force] = dict(h2=h2.detach().float(), e2=e2.detach().float(), dh=hg.grad.float(), de=eg.grad.float(),
                           path=lib.egt_last_path(),
                           **{f'g_{k}': blk.grad_view(k).clone() for k in blk.layout})
    lib.egt_debug_force_staged(0)
    a, b = outs[1], outs[0]
    line = [f'N={N} B={B} train={training} path={b["path"]}']
    for k in a:
        if k == 'path':
            continue
        ref = a[k]
        err = float((b[k] - ref).abs().max())
        sc = float(ref.abs().max())
        flag = '' if err <= 3e-2 * max(sc, 1e-3) else '  <-- BAD'
        line.append(f'  {k:28s} err {err:.3e} / scale {sc:.3e}{flag}')
    print('\n'.join(line), flush=True)


if __name__ == '__main__':
    cases = [(128, 4), (75, 3), (37, 3), (190, 2), (9, 2), (1, 2), (257, 2), (16, 2), (17, 2)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in c.split(',')) for c in sys.argv[1:]]
    for N, B in cases:
        for tr, rmp in ((False, 0.0), (True, 0.1)):
            run(N, B, tr, rmp)
