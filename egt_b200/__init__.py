"""egt_b200 -- B200-native (sm_100a) implementation of the EGT edge-augmented attention block.

Public surface (mirrors the reference's layer interface for this path only):
    EGT        drop-in for lib/models/egt_layers.py:EGT
    EGTBlock   one attention block (h, e, mask) -> (h', e') with the reference's weight names
    EGTFFN     feed-forward half of a layer for one channel ("next" row: ffn_block of the reference)
    EGTLayer / EGTStack   one layer / the L-layer loop (graph_xformer_model_base.py:335-341) on ONE flat parameter
    DataParallelBlock / allreduce_flat_grads   one NCCL all-reduce per step on the flat gradient
"""
from .layers import EGT, EGTBlock, EGTFFN, EGTLayer, EGTStack
from .ops import AttnSpec, BlockSpec, egt_attention, egt_block, egt_ffn
from .dp import allreduce_flat_grads, shard_batch

__all__ = ['EGT', 'EGTBlock', 'EGTFFN', 'EGTLayer', 'EGTStack', 'AttnSpec', 'BlockSpec', 'egt_attention', 'egt_block', 'egt_ffn',
           'allreduce_flat_grads', 'shard_batch']
