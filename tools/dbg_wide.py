"""Bring-up aid for the width-generic fused kernels: runs one block forward (+ backward) per width on the fused and on
the staged path and prints error statistics against the oracle instead of stopping at the first failure.

usage: python tools/dbg_wide.py [C5 C1 C3] [N] [B]"""
import sys

import torch

sys.path.insert(0, '.')
from oracle import egt_oracle as O                      # noqa: E402
from tests.test_wide_gpu import _case, _fwd_bwd        # noqa: E402


def stats(name, got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    err = (got - ref).abs()
    bad = err > 1e-2 * max(1.0, float(ref.abs().max())) + 1e-2 * ref.abs()
    idx = err.flatten().argmax()
    where = tuple(int(v) for v in torch.unravel_index(idx, err.shape))
    print(f'  {name:10s} max|ref| {float(ref.abs().max()):9.3e}  max err {float(err.max()):9.3e} at {where}  '
          f'bad {int(bad.sum())}/{bad.numel()}  nan {int(torch.isnan(got).sum())}', flush=True)


def main():
    widths = [a for a in sys.argv[1:] if a in ('C5', 'C1', 'C3')] or ['C5', 'C1', 'C3']
    nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
    N = nums[0] if nums else 37
    B = nums[1] if len(nums) > 1 else 2
    for width in widths:
        for training, rmp in ((False, 0.), (True, 0.1)):
            cfg, params, h, e, mask = _case(width, N, B, training, rmp)
            for force in (1, 0):
                print(f'{width} N={N} B={B} training={training} force_staged={force}', flush=True)
                try:
                    blk, (h2, e2, gin), (h2r, e2r, rin, pr), paths = _fwd_bwd(cfg, params, h, e, mask, training, force)
                except Exception as ex:   # keep going: the next configuration may still tell something
                    print('  FAILED:', type(ex).__name__, str(ex)[:300], flush=True)
                    continue
                print('  paths', paths)
                stats("h'", h2, h2r)
                stats("e'", e2, e2r)
                stats('dh', gin[0], rin[0])
                stats('de', gin[1], rin[1])
                blk.flat.grad = gin[2]
                for (name, _), gr in zip(pr.items(), rin[2:]):
                    stats(name[-10:], blk.grad_view(name.replace('/', '_')), gr)


if __name__ == '__main__':
    main()
