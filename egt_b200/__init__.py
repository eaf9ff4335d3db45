"""egt_b200 -- B200-native (sm_100a) implementation of the EGT edge-augmented attention block.

Public surface (mirrors the reference's layer interface for this path only):
    EGT        drop-in for lib/models/egt_layers.py:EGT
    EGTBlock   one attention block (h, e, mask) -> (h', e') with the reference's weight names
    DataParallelBlock / allreduce_flat_grads   one NCCL all-reduce per step on the flat gradient
"""
from .layers import EGT, EGTBlock
from .ops import AttnSpec, BlockSpec, egt_attention, egt_block
from .dp import allreduce_flat_grads, shard_batch

__all__ = ['EGT', 'EGTBlock', 'AttnSpec', 'BlockSpec', 'egt_attention', 'egt_block',
           'allreduce_flat_grads', 'shard_batch']
