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


class _PeerAllReduce:
    """One-shot all-reduce over NVLink peer memory (csrc/peer_allreduce.cu).  The payload is a few tens of KB, so this
    is a latency play.  Default protocol ``push``: every rank stores its gradient, with the call number inside every
    8-byte word, straight into a receive slot of every peer and adds the slots of its own buffer -- no fence, no flag
    round trip, no remote load.  ``EGT_PEER_PROTOCOL=pull`` selects the first version (publish, flag exchange, remote
    loads), which measured 37 us per call on 8 GPUs against NCCL's 29."""

    def __init__(self, numel: int, device: torch.device):
        import ctypes as C
        import os
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self._C, self._L = C, L
        self.numel = numel
        world = dist.get_world_size()
        self.push = os.environ.get('EGT_PEER_PROTOCOL', 'push') != 'pull' and numel % 2 == 0
        lib = L.load()
        self.buf_floats = int(lib.egt_peer_allreduce_push_floats(numel, world)) if self.push else 2 * numel
        self.buf = symm.empty(self.buf_floats, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        assert self.hdl.signal_pad_size >= 4 * (32 * 8 + 32), 'signal pad too small'
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier()

    def __call__(self, grad: torch.Tensor):
        C, L = self._C, self._L
        lib = L.load()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.push:
            L.check(lib.egt_peer_allreduce_push(C.c_void_p(self.hdl.buffer_ptrs_dev), C.c_void_p(self.hdl.signal_pad_ptrs_dev),
                                                C.c_void_p(grad.data_ptr()), self.numel, self.buf_floats, self.hdl.rank,
                                                self.hdl.world_size, stream))
        else:
            L.check(lib.egt_peer_allreduce(C.c_void_p(self.hdl.buffer_ptrs_dev), C.c_void_p(self.hdl.signal_pad_ptrs_dev),
                                           C.c_void_p(grad.data_ptr()), self.numel, self.hdl.rank, self.hdl.world_size, stream))


_peer_cache = {}


def _peer_allreduce_for(grad: torch.Tensor):
    """The peer-memory all-reduce for this gradient buffer, or None when it cannot be used (CPU / gloo,
    more than 8 ranks, unaligned size, symmetric memory unavailable, or EGT_PEER_ALLREDUCE=0)."""
    import os
    if os.environ.get('EGT_PEER_ALLREDUCE', '1') == '0' or not grad.is_cuda or not grad.is_contiguous():
        return None
    if grad.dtype != torch.float32 or grad.numel() % 4 or grad.data_ptr() % 16 or dist.get_world_size() > 8:
        return None
    key = (grad.numel(), grad.device.index)
    if key not in _peer_cache:
        try:
            _peer_cache[key] = _PeerAllReduce(grad.numel(), grad.device)
        except Exception as ex:   # no symmetric memory on this system: the NCCL collective is always available
            import warnings
            warnings.warn(f'egt_b200: peer-memory all-reduce unavailable ({type(ex).__name__}: {ex}); using NCCL')
            _peer_cache[key] = None
    return _peer_cache[key]


def allreduce_flat_grads(blocks: Iterable, async_op: bool = False):
    """Sum ``block.flat.grad`` over ranks.  With one block this is literally one collective on one
    buffer; several blocks are coalesced into one flat tensor first so it stays one collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    blocks = list(blocks)
    if not blocks:
        return None
    # every rank must take part with the SAME buffers (the peer kernel counts its calls per buffer): a block
    # that got no gradient on this rank contributes zeros instead of silently dropping out of the collective
    for b in blocks:
        if b.flat.grad is None:
            b.flat.grad = torch.zeros_like(b.flat)
    grads = [b.flat.grad for b in blocks]
    if len(grads) == 1:
        peer = None if async_op else _peer_allreduce_for(grads[0])
        if peer is not None:
            peer(grads[0])
            return None
        return dist.all_reduce(grads[0], op=dist.ReduceOp.SUM, async_op=async_op)
    flat = torch.cat([g.reshape(-1) for g in grads])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=False)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return work
