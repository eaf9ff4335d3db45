"""Host-side mirror of the reference's layer interface for the hot path.

``EGT``       <-> lib/models/egt_layers.py:4-217   (same constructor kwargs, same positional input list)
``EGTBlock``  <-> edge_update_{none,bias,residual} + mha_block of
                  lib/models/graph_xformer_model_base.py:106-223 for one layer ``tag``; owns the
                  weights under the reference's layer names.
"""
import math
from typing import Dict, Optional

import torch
from torch import nn

from . import ops

_KERAS_NAMES = {   # flat-buffer field -> (reference layer name stem, keras weight name)
    'norm_mha_gamma': ('norm_mha', 'gamma'), 'norm_mha_beta': ('norm_mha', 'beta'),
    'dense_qkv_kernel': ('dense_qkv', 'kernel'), 'dense_qkv_bias': ('dense_qkv', 'bias'),
    'dense_mha_kernel': ('dense_mha', 'kernel'), 'dense_mha_bias': ('dense_mha', 'bias'),
    'norm_edge_gamma': ('norm_edge', 'gamma'), 'norm_edge_beta': ('norm_edge', 'beta'),
    'attention_gates_kernel': ('attention_gates', 'kernel'), 'attention_gates_bias': ('attention_gates', 'bias'),
    'dense_edge_b_kernel': ('dense_edge_b', 'kernel'), 'dense_edge_b_bias': ('dense_edge_b', 'bias'),
    'dense_edge_r_kernel': ('dense_edge_r', 'kernel'), 'dense_edge_r_bias': ('dense_edge_r', 'bias'),
}


class _RngState:
    """seed/offset bookkeeping for the in-kernel counter RNG (random key mask, dropout)."""

    def __init__(self, seed=0):
        self.seed = seed
        self.offset = 0

    def next(self):
        self.offset += 1
        return self.seed, self.offset


def _rng_args(module, training, live):
    """(seed, offset, offset_dev) of one call.  Eager calls advance the host offset.  A call that is being captured into
    a CUDA graph with live RNG (random key mask / attention dropout) also takes a snapshot of the module's DEVICE counter
    and bumps it, both as captured work: the kernels add the snapshot to the (frozen) host offset, so every replay of the
    graph draws new noise, and the backward of the call reuses the forward's snapshot."""
    if not training:
        return 0, 0, None
    seed, offset = module.rng.next()
    if live and module._rng_counter.is_cuda and torch.cuda.is_current_stream_capturing():
        return seed, offset, ops.graph_safe_offset(module._rng_counter)
    return seed, offset, None


class EGT(nn.Module):
    """Drop-in for the reference Keras layer ``EGT`` (egt_layers.py:4-40).  No weights.

    ``forward([QKV, E?, G?, M?], mask=None, training=None) -> (V_att, H_hat, A_tild)``;
    ``A_tild`` is ``None`` unless ``return_attn=True`` (only Analysis taps consume it,
    graph_xformer_model_base.py:134)."""

    def __init__(self, num_heads=8, clip_logits_value=(-5, 5), scale_degree=False, scaler_type='log',
                 edge_input=True, gate_input=True, attn_mask=False, num_virtual_nodes=0,
                 random_mask_prob=0., attn_dropout=0., return_attn=False, seed=0):
        super().__init__()
        clip = None if clip_logits_value is None else tuple(float(v) for v in clip_logits_value)
        self.spec = ops.AttnSpec(num_heads=num_heads, clip_logits_value=clip, scale_degree=scale_degree,
                                 scaler_type=scaler_type, edge_input=edge_input, gate_input=gate_input,
                                 attn_mask=attn_mask, num_virtual_nodes=num_virtual_nodes,
                                 random_mask_prob=random_mask_prob, attn_dropout=attn_dropout)
        self.spec.validate()
        self.return_attn = return_attn
        self.rng = _RngState(seed)
        self.register_buffer('_rng_counter', torch.zeros(1, dtype=torch.int64), persistent=False)   # see _rng_args

    def get_config(self):
        """Same keys as egt_layers.py:42-55, plus ``attn_dropout`` which the reference forgets."""
        s = self.spec
        return dict(num_heads=s.num_heads, clip_logits_value=s.clip_logits_value, scale_degree=s.scale_degree,
                    scaler_type=s.scaler_type, edge_input=s.edge_input, gate_input=s.gate_input,
                    attn_mask=s.attn_mask, num_virtual_nodes=s.num_virtual_nodes,
                    random_mask_prob=s.random_mask_prob, attn_dropout=s.attn_dropout)

    def compute_mask(self, inputs, mask):                                    # egt_layers.py:215-217
        return [mask[0] if isinstance(mask, (list, tuple)) else mask, None, None]

    def forward(self, inputs, mask=None, training=None):
        if training is None:                                                 # egt_layers.py:58-59
            training = self.training
        seed, offset, offset_dev = _rng_args(self, training, self.spec.random_mask_prob > 0 or self.spec.attn_dropout > 0)
        return ops.egt_attention(inputs, mask, training, spec=self.spec, seed=seed, offset=offset,
                                 return_attn=self.return_attn, offset_dev=offset_dev)


class EGTBlock(nn.Module):
    """One attention block ``(h, e, mask) -> (h', e')``.

    All weights live in ONE flat float32 parameter (``self.flat``) so that data-parallel training
    needs a single all-reduce on ``self.flat.grad``; named views follow the reference's layer names
    (``norm_mha``, ``dense_qkv``, ``dense_mha``, ``norm_edge``, ``attention_gates``,
    ``dense_edge_b``, ``dense_edge_r``)."""

    def __init__(self, tag='00', seed=0, **kwargs):
        super().__init__()
        if 'clip_logits_value' in kwargs and kwargs['clip_logits_value'] is not None:
            kwargs['clip_logits_value'] = tuple(float(v) for v in kwargs['clip_logits_value'])
        self.spec = ops.BlockSpec(**kwargs)
        self.spec.validate()
        self.tag = tag
        total, self.layout = ops.param_layout(self.spec)
        self.flat = nn.Parameter(torch.zeros(total, dtype=torch.float32))
        self.rng = _RngState(seed)
        self.register_buffer('_rng_counter', torch.zeros(1, dtype=torch.int64), persistent=False)   # see _rng_args
        self.reset_parameters()

    # -- parameters ---------------------------------------------------------------------
    def view(self, field: str) -> torch.Tensor:
        off, shape = self.layout[field]
        return self.flat.data[off:off + math.prod(shape)].view(shape)

    def grad_view(self, field: str) -> Optional[torch.Tensor]:
        if self.flat.grad is None:
            return None
        off, shape = self.layout[field]
        return self.flat.grad[off:off + math.prod(shape)].view(shape)

    def reset_parameters(self, seed=1234):
        """Keras defaults: Dense = glorot_uniform kernel / zero bias; LayerNormalization = ones / zeros."""
        g = torch.Generator().manual_seed(seed)
        for f in self.layout:
            v = self.view(f)
            if f.endswith('kernel'):
                lim = math.sqrt(6.0 / (v.shape[0] + v.shape[1]))
                v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * lim)
            elif f.endswith('gamma'):
                v.fill_(1.)
            else:
                v.zero_()

    def keras_weights(self) -> Dict[str, torch.Tensor]:
        """{'<layer>_{tag}/<weight>': tensor} with Keras layouts (Dense kernel [in,out])."""
        return {f'{_KERAS_NAMES[f][0]}_{self.tag}/{_KERAS_NAMES[f][1]}': self.view(f).clone() for f in self.layout}

    def load_keras_weights(self, weights: Dict[str, torch.Tensor], strict=True):
        """Load by reference layer name (what ``load_weights(by_name=True)`` does,
        lib/training/training_base.py:362).  Accepts keys with or without the ``_{tag}`` suffix."""
        for f in self.layout:
            stem, wn = _KERAS_NAMES[f]
            for key in (f'{stem}_{self.tag}/{wn}', f'{stem}/{wn}'):
                if key in weights:
                    w = torch.as_tensor(weights[key], dtype=torch.float32)
                    assert tuple(w.shape) == tuple(self.layout[f][1]), f'{key}: {tuple(w.shape)} != {self.layout[f][1]}'
                    self.view(f).copy_(w)
                    break
            else:
                if strict:
                    raise KeyError(f'missing weight for {stem}_{self.tag}/{wn}')

    # -- call ---------------------------------------------------------------------------
    def forward(self, h, e, mask=None, edge_mask=None, training=None):
        if training is None:
            training = self.training
        seed, offset, offset_dev = _rng_args(self, training, self.spec.random_mask_prob > 0 or self.spec.attn_dropout > 0)
        return ops.egt_block(h, e, mask, self.flat, self.spec, self.layout, edge_mask=edge_mask,
                             training=training, seed=seed, offset=offset, offset_dev=offset_dev)


class EGTFFN(nn.Module):
    """Feed-forward half of a layer for ONE channel (node: ``width = model_width``; edge: ``width =
    edge_width``): ``y = x + Dense_w(act(Dense_round(w*ffn_multiplier)(LayerNorm(x))))`` --
    ``ffnlr1 / ffnact / ffnlr2`` of graph_xformer_model_base.py:229-258 as ``ffn_block`` (:309-324) chains
    them.  Weights live in one flat float32 parameter; names follow the reference
    (``norm_fnn_{channel}_{tag}``, ``fnn_lr1_{channel}_{tag}``, ``fnn_lr2_{channel}_{tag}``)."""

    _NAMES = {'norm_gamma': ('norm_fnn', 'gamma'), 'norm_beta': ('norm_fnn', 'beta'),
              'lr1_kernel': ('fnn_lr1', 'kernel'), 'lr1_bias': ('fnn_lr1', 'bias'),
              'lr2_kernel': ('fnn_lr2', 'kernel'), 'lr2_bias': ('fnn_lr2', 'bias')}

    def __init__(self, width, channel='node', tag='00', ffn_multiplier=2., activation='elu', ln_eps=1e-3):
        super().__init__()
        assert channel in ('node', 'edge')
        self.width, self.hidden = int(width), int(round(width * ffn_multiplier))   # graph_xformer_model_base.py:236
        self.channel, self.tag, self.activation, self.ln_eps = channel, tag, activation, ln_eps
        total, self.layout = ops.ffn_layout(self.width, self.hidden)
        self.flat = nn.Parameter(torch.zeros(total, dtype=torch.float32))
        self.reset_parameters()

    def view(self, field):
        off, shape = self.layout[field]
        return self.flat.data[off:off + math.prod(shape)].view(shape)

    def grad_view(self, field):
        if self.flat.grad is None:
            return None
        off, shape = self.layout[field]
        return self.flat.grad[off:off + math.prod(shape)].view(shape)

    def reset_parameters(self, seed=4321):
        g = torch.Generator().manual_seed(seed)
        for f in self.layout:
            v = self.view(f)
            if f.endswith('kernel'):
                lim = math.sqrt(6.0 / (v.shape[0] + v.shape[1]))
                v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * lim)
            elif f.endswith('gamma'):
                v.fill_(1.)
            else:
                v.zero_()

    def load_keras_weights(self, weights: Dict[str, torch.Tensor], strict=True):
        """Keys ``'{stem}_{channel}_{tag}/{weight}'`` as in the reference checkpoint, or the oracle's
        ``'ffn_{channel}/{norm|lr1|lr2}/{weight}'``."""
        short = {'norm_fnn': 'norm', 'fnn_lr1': 'lr1', 'fnn_lr2': 'lr2'}
        for f in self.layout:
            stem, wn = self._NAMES[f]
            for key in (f'{stem}_{self.channel}_{self.tag}/{wn}', f'ffn_{self.channel}/{short[stem]}/{wn}'):
                if key in weights:
                    w = torch.as_tensor(weights[key], dtype=torch.float32)
                    assert tuple(w.shape) == tuple(self.layout[f][1]), f'{key}: {tuple(w.shape)} != {self.layout[f][1]}'
                    self.view(f).copy_(w)
                    break
            else:
                if strict:
                    raise KeyError(f'missing weight for {stem}_{self.channel}_{self.tag}/{wn}')

    def forward(self, x):
        return ops.egt_ffn(x, self.flat, self.width, self.hidden, self.layout, self.activation, self.ln_eps)


class EGTLayer(nn.Module):
    """One full EGT layer as the reference's loop body builds it (graph_xformer_model_base.py:335-341):
    ``edge_update(tag, h, e)`` (the attention block) followed by ``ffn_block(tag, h, e)`` (node FFN, and the
    edge FFN when the edge channel is residual / constrained)."""

    def __init__(self, tag='00', ffn_multiplier=2., activation='elu', seed=0, **block_kwargs):
        super().__init__()
        self.block = EGTBlock(tag=tag, seed=seed, **block_kwargs)
        sp = self.block.spec
        self.ffn_node = EGTFFN(sp.model_width, 'node', tag, ffn_multiplier, activation, sp.ln_eps)
        self.ffn_edge = EGTFFN(sp.edge_width, 'edge', tag, ffn_multiplier, activation, sp.ln_eps) if sp.is_residual else None

    def load_keras_weights(self, weights, strict=True):
        self.block.load_keras_weights({k: v for k, v in weights.items() if not k.startswith('ffn') and 'fnn' not in k}, strict)
        self.ffn_node.load_keras_weights(weights, strict)
        if self.ffn_edge is not None:
            self.ffn_edge.load_keras_weights(weights, strict)

    def forward(self, h, e, mask=None, edge_mask=None, training=None):
        h, e = self.block(h, e, mask, edge_mask=edge_mask, training=training)
        h = self.ffn_node(h)
        if self.ffn_edge is not None:
            e = self.ffn_edge(e)
        return h, e


class EGTStack(nn.Module):
    """``L`` EGT layers as the reference's layer loop chains them (graph_xformer_model_base.py:335-341:
    ``for ii in range(model_height): h, e = edge_update(tag, h, e); h, e = ffn_block(tag, h, e)``) with EVERY weight
    of every layer in ONE flat float32 parameter, so that data-parallel training is a single all-reduce on
    ``self.flat.grad`` (``egt_b200.allreduce_flat_grads([stack])``; MirroredStrategy semantics,
    lib/training/training_base.py:230-238).  Layer ``ii`` carries the reference's tag ``f'{ii:0>2d}'`` (weights load
    by the reference's names); ``ffn=False`` builds the attention blocks only."""

    def __init__(self, model_height, ffn=True, ffn_multiplier=2., activation='elu', seed=0, **block_kwargs):
        super().__init__()
        self.layers = nn.ModuleList()
        for ii in range(model_height):
            tag = f'{ii:0>2d}'
            self.layers.append(EGTLayer(tag=tag, ffn_multiplier=ffn_multiplier, activation=activation, seed=seed + ii, **block_kwargs)
                               if ffn else EGTBlock(tag=tag, seed=seed + ii, **block_kwargs))
        # every module of every layer that owns a flat parameter, in execution order
        self._owners = []
        for layer in self.layers:
            mods = [layer] if isinstance(layer, EGTBlock) else [layer.block, layer.ffn_node] + ([layer.ffn_edge] if layer.ffn_edge is not None else [])
            self._owners.extend(mods)
        self._spans, total = [], 0
        for m in self._owners:
            n = m.flat.numel()
            self._spans.append((total, n))
            total += (n + 3) // 4 * 4                    # 16-byte aligned slices (the kernels use vector accesses)
        flat = torch.zeros(total, dtype=torch.float32)
        for m, (off, n) in zip(self._owners, self._spans):
            flat[off:off + n].copy_(m.flat.data)
            del m._parameters['flat']                    # the module keeps working on a VIEW of the stack's parameter
            m.flat = flat[off:off + n]
        self.flat = nn.Parameter(flat)
        self._bind()

    def _bind(self):
        """Re-create the per-module views (after ``.to(device)`` or before every forward: autograd must see them as
        slices of the ONE parameter)."""
        # drop the previous views first: they keep the autograd node of self.flat from an earlier call alive (on
        # whatever stream that call ran), which breaks CUDA-graph capture of a later call
        for m in self._owners:
            m.flat = None
        for m, (off, n) in zip(self._owners, self._spans):
            m.flat = self.flat[off:off + n]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._bind()
        return out

    def load_keras_weights(self, weights, strict=True):
        with torch.no_grad():
            self._bind()
            for layer in self.layers:
                layer.load_keras_weights(weights, strict)

    def grad_view(self, index, field):
        """Gradient of ``field`` of the ``index``-th flat-owning module (execution order), a view into ``self.flat.grad``."""
        if self.flat.grad is None:
            return None
        m, (off, n) = self._owners[index], self._spans[index]
        o, shape = m.layout[field]
        return self.flat.grad[off + o:off + o + math.prod(shape)].view(shape)

    def forward(self, h, e, mask=None, edge_mask=None, training=None):
        self._bind()
        if training is None:
            training = self.training
        for layer in self.layers:
            h, e = layer(h, e, mask, edge_mask=edge_mask, training=training)
        return h, e
