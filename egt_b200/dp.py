"""Data-parallel plumbing: batches shard by graph, weight gradients are summed with ONE all-reduce.

Mirrors what ``tf.distribute.MirroredStrategy`` does for the reference
(lib/training/training_base.py:230-238): ``batch_size`` is the GLOBAL batch, each replica gets an
equal slice on the batch axis, per-replica gradients of ``local_loss_sum / global_batch`` are
SUMMED.  No other collective exists on this path (graphs are independent)."""
from typing import Iterable

import torch
import torch.distributed as dist


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Slice the global batch axis for this rank (even split; the reference drops nothing because
    Keras splits each dataset batch across replicas)."""
    B = t.shape[0]
    if B % world:
        raise ValueError(f'global batch {B} is not divisible by world size {world}')
    per = B // world
    return t[rank * per:(rank + 1) * per]


def allreduce_flat_grads(blocks: Iterable, async_op: bool = False):
    """Sum ``block.flat.grad`` over ranks.  With one block this is literally one collective on one
    buffer; several blocks are coalesced into one flat tensor first so it stays one collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    grads = [b.flat.grad for b in blocks if b.flat.grad is not None]
    if not grads:
        return None
    if len(grads) == 1:
        return dist.all_reduce(grads[0], op=dist.ReduceOp.SUM, async_op=async_op)
    flat = torch.cat([g.reshape(-1) for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=False)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return work
