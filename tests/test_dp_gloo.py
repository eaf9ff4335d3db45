"""CPU, world_size 2, gloo: the data-parallel plumbing (batch sharding + ONE all-reduce of the flat
gradient) gives the same gradient as a single process on the global batch.
The per-rank gradients come from the CPU oracle here -- this tests the host-side logic only."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat_grad_oracle(blk, h, e, mask, cfg, global_B):
    from oracle import egt_oracle as O
    params = {k.replace('_00/', '/'): v.double().requires_grad_(True) for k, v in blk.keras_weights().items()}
    h2, e2 = O.egt_block(h.double(), e.double(), mask, params, cfg)
    loss = (h2.pow(2).sum() + e2.pow(2).sum()) / global_B           # local_loss_sum / global_batch
    grads = torch.autograd.grad(loss, list(params.values()))
    flat = torch.zeros_like(blk.flat.data, dtype=torch.float64)
    for (name, _), g in zip(params.items(), grads):
        off, shape = blk.layout[name.replace('/', '_')]
        flat[off:off + g.numel()] = g.reshape(-1)
    return flat


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import egt_b200
    from oracle import egt_oracle as O
    cfg = O.BlockConfig(model_width=16, edge_width=8, num_heads=4, scale_degree=True)
    blk = egt_b200.EGTBlock(model_width=16, edge_width=8, num_heads=4, scale_degree=True)
    h, e, mask = O.synthetic_batch(4, 6, 16, 8, ragged=True)
    hs, es, ms = (egt_b200.shard_batch(t, rank, world) for t in (h, e, mask))
    assert hs.shape[0] == 2
    blk.flat.grad = _flat_grad_oracle(blk, hs, es, ms, cfg, 4).float()
    egt_b200.allreduce_flat_grads([blk])
    if rank == 0:
        ref = _flat_grad_oracle(blk, h, e, mask, cfg, 4).float()
        out.put((blk.flat.grad.clone(), ref))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, ref = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)


def test_shard_batch_rejects_uneven():
    import egt_b200
    with pytest.raises(ValueError):
        egt_b200.shard_batch(torch.zeros(5, 2), 0, 2)
