"""CPU oracle for the EGT edge-augmented attention block  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (PyTorch CPU tensors, fp32 or fp64, eager, every
``[B,N,N,h]`` intermediate materialised exactly as the TF graph does) of the
reference's hot path:

  * ``lib/models/egt_layers.py:57-143``   (``EGT.call_gated``)
  * ``lib/models/egt_layers.py:145-213``  (``EGT.call_ungated``)
  * ``lib/models/graph_xformer_model_base.py:106-145``  (``mha_block``)
  * ``lib/models/graph_xformer_model_base.py:149-162``  (``edge_channel_contrib``)
  * ``lib/models/graph_xformer_model_base.py:164-223``  (``edge_update_{none,bias,residual}``)
  * ``lib/models/graph_xformer_model_base.py:229-258,309-324`` (``ffn_block``, the "next" row)

It is NOT the product: nothing under ``egt_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

Parity status: the reference ships no tests, fixtures or golden vectors and TensorFlow
is not installable in this image, so parity with the *real TF runtime* is UNPINNED.
What pins this restatement instead (see ``oracle/make_golden.py``):
  1. the reference's own source files (``egt_layers.py`` and the ``mha_block`` /
     ``edge_update_*`` closures of ``graph_xformer_model_base.py``) are imported
     unmodified from ``/root/reference`` and executed on top of a small TF/Keras *shim*
     (``oracle/tf_shim``) that maps each TF op they call onto its documented semantics;
     their outputs are committed as ``tests/golden/*.npz`` and this restatement must
     reproduce them;
  2. ``torch.autograd.gradcheck`` + a hand-derived closed-form backward
     (``egt_block_backward``) checked against autograd in fp64.

Keras defaults relied upon (inferred, not in the reference tree): ``LayerNormalization``
axis=-1, epsilon=1e-3; ``Dense`` kernel is ``[in,out]`` applied to the last axis;
``tf.nn.dropout`` scales kept values by ``1/(1-rate)``; ``tf.clip_by_value`` passes
gradient where ``lo <= x <= hi``; softmax subtracts the row max.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import torch

NEG_INF_MASK = 1e9          # egt_layers.py:92,99,106 use (x-1)*1e9 and -1e9
LN_EPS = 1e-3               # keras.layers.LayerNormalization default epsilon (inferred)


# --------------------------------------------------------------------------------------
# Keras primitives the block is assembled from
# --------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
               eps: float = LN_EPS) -> torch.Tensor:
    """keras.layers.LayerNormalization(axis=-1) (graph_xformer_model_base.py:97-100,109,195)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def dense(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """keras.layers.Dense on the last axis; kernel is [in, out]."""
    y = x @ kernel
    if bias is not None:
        y = y + bias
    return y


def edge_activation_fn(name: Optional[str]):
    """edge_channel_contrib's activation (graph_xformer_model_base.py:149-162).

    'lreluX' -> LeakyReLU(alpha=X/10) where X is the LAST character (:150-152)."""
    if name is None:
        return lambda v: v
    low = name.lower()
    if low.startswith('lrelu'):
        alpha = float(name[-1]) / 10
        return lambda v: torch.where(v >= 0, v, alpha * v)
    table = dict(relu=torch.relu, elu=torch.nn.functional.elu, tanh=torch.tanh,
                 sigmoid=torch.sigmoid, linear=lambda v: v)
    if low not in table:
        raise ValueError(f'unsupported edge_activation {name!r}')
    return table[low]


# --------------------------------------------------------------------------------------
# EGT layer  (egt_layers.py)
# --------------------------------------------------------------------------------------
def egt_layer(inputs: Sequence[torch.Tensor], mask: Optional[torch.Tensor] = None,
              training: bool = False, *,
              num_heads: int = 8,
              clip_logits_value: Optional[Sequence[float]] = (-5., 5.),
              scale_degree: bool = False,
              scaler_type: str = 'log',
              edge_input: bool = True,
              gate_input: bool = True,
              attn_mask: bool = False,
              num_virtual_nodes: int = 0,
              random_mask_prob: float = 0.0,
              attn_dropout: float = 0.0,
              uniform_noise: Optional[torch.Tensor] = None,
              dropout_noise: Optional[torch.Tensor] = None,
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``EGT.call`` : ([QKV, E?, G?, M?], mask, training) -> (V_att, H_hat, A_tild).

    Follows egt_layers.py:57-143 (gated) / :145-213 (ungated) statement by statement.
    ``uniform_noise`` / ``dropout_noise`` are the tensors ``tf.random.uniform`` /
    ``tf.nn.dropout`` would draw (``[B,N,N,h]`` in [0,1)); injected so results are
    reproducible.  They are only consulted when the reference would draw them.
    """
    # constructor checks (egt_layers.py:20-24)
    if scale_degree and not gate_input:
        raise ValueError('scale_degree requires gate_input')
    if scaler_type not in ('log', 'linear'):
        raise ValueError('scaler_type must be log or linear')

    inputs = list(inputs)
    QKV = inputs.pop(0)                                   # :62  b,l,3dh
    E = inputs.pop(0) if edge_input else None             # :63
    G = inputs.pop(0) if gate_input else None             # :64
    M = inputs.pop(0) if attn_mask else None              # :65
    if isinstance(mask, (list, tuple)):                   # :66
        mask = mask[0]

    B, N, C = QKV.shape
    assert C % (num_heads * 3) == 0                       # :70
    dot_dim = C // (num_heads * 3)                        # :71
    QKV5 = QKV.reshape(B, N, 3, dot_dim, num_heads)       # :73-75   b,l,3,d,h
    Q, K, V = QKV5.unbind(dim=2)                          # :76      b,l,d,h

    A_hat = torch.einsum('bldh,bmdh->blmh', Q, K) * (dot_dim ** -0.5)      # :79
    if clip_logits_value is not None:                     # :81-82
        A_hat = torch.clamp(A_hat, clip_logits_value[0], clip_logits_value[1])

    H_hat = A_hat                                         # :85
    if edge_input:
        H_hat = H_hat + E                                 # :86

    H_hat_ = H_hat                                        # :89
    G_ = G                                                # :90
    if mask is not None:                                  # :91-94
        mask_ = (mask[:, None, :, None].to(H_hat.dtype) - 1) * NEG_INF_MASK
        H_hat_ = H_hat_ + mask_
        if gate_input:
            G_ = G_ + mask_
    if attn_mask:                                         # :96-101
        M_ = (M.to(H_hat.dtype) - 1) * NEG_INF_MASK
        H_hat_ = H_hat_ + M_
        if gate_input:
            G_ = G_ + M_
    if random_mask_prob > 0.0 and training:               # :103-108
        assert uniform_noise is not None, 'inject the tf.random.uniform draw'
        random_mask_ = torch.where(uniform_noise.to(H_hat.dtype) < random_mask_prob,
                                   torch.full_like(H_hat, -NEG_INF_MASK),
                                   torch.zeros_like(H_hat))
        H_hat_ = H_hat_ + random_mask_
        if gate_input:
            G_ = G_ + random_mask_

    A_tild = torch.softmax(H_hat_, dim=2)                 # :111
    gates = None
    if gate_input:
        gates = torch.sigmoid(G_)                         # :112
        A_tild = A_tild * gates                           # :113

    if attn_dropout > 0.0 and training:                   # :116-117  tf.nn.dropout
        assert dropout_noise is not None, 'inject the tf.nn.dropout draw'
        keep = (dropout_noise >= attn_dropout).to(A_tild.dtype)
        A_tild = A_tild * keep / (1.0 - attn_dropout)

    V_att = torch.einsum('blmh,bmdh->bldh', A_tild, V)    # :120

    if scale_degree:                                      # :123-136
        degrees = gates.sum(dim=2, keepdim=True)          # b,l,1,h
        if scaler_type == 'log':
            degree_scalers = torch.log(1 + degrees)
        else:
            degree_scalers = degrees
        if num_virtual_nodes > 0:
            non_vn = degree_scalers[:, num_virtual_nodes:]
            ones = torch.ones_like(degree_scalers[:, :num_virtual_nodes])
            degree_scalers = torch.cat([ones, non_vn], dim=1)
        V_att = V_att * degree_scalers

    V_att = V_att.reshape(B, N, dot_dim * num_heads)      # :139-141
    return V_att, H_hat, A_tild                           # :143


# --------------------------------------------------------------------------------------
# attention block = edge_update_* around mha_block  (graph_xformer_model_base.py)
# --------------------------------------------------------------------------------------
@dataclass
class BlockConfig:
    """The GraphTransformerBase constructor arguments that reach the attention block
    (graph_xformer_model_base.py:17-45)."""
    model_width: int = 128
    edge_width: int = 32
    num_heads: int = 8
    gate_attention: bool = True
    node_dropout: float = 0.0
    edge_dropout: float = 0.0
    add_n_norm: bool = False
    clip_logits_value: Optional[Sequence[float]] = (-5., 5.)
    edge_activation: Optional[str] = None
    edge_channel_type: str = 'residual'      # none | bias | residual | constrained
    ffn_multiplier: float = 2.0
    activation: str = 'elu'
    scale_degree: bool = False
    scaler_type: str = 'log'
    num_virtual_nodes: int = 0
    random_mask_prob: float = 0.0
    attn_dropout: float = 0.0

    def __post_init__(self):
        if not self.gate_attention and self.scale_degree:      # :46-47
            raise ValueError('scale_degree only works with gate_attention')


PARAM_SHAPES = {
    # name: lambda cfg -> shape           (layer names of Appendix D without the tag)
    'norm_mha/gamma':        lambda c: (c.model_width,),
    'norm_mha/beta':         lambda c: (c.model_width,),
    'dense_qkv/kernel':      lambda c: (c.model_width, 3 * c.model_width),
    'dense_qkv/bias':        lambda c: (3 * c.model_width,),
    'dense_mha/kernel':      lambda c: (c.model_width, c.model_width),
    'dense_mha/bias':        lambda c: (c.model_width,),
    'norm_edge/gamma':       lambda c: (c.edge_width,),
    'norm_edge/beta':        lambda c: (c.edge_width,),
    'attention_gates/kernel': lambda c: (c.edge_width, c.num_heads),
    'attention_gates/bias':   lambda c: (c.num_heads,),
    'dense_edge_b/kernel':   lambda c: (c.edge_width, c.num_heads),
    'dense_edge_b/bias':     lambda c: (c.num_heads,),
    'dense_edge_r/kernel':   lambda c: (c.num_heads, c.edge_width),
    'dense_edge_r/bias':     lambda c: (c.edge_width,),
}


def block_param_names(cfg: BlockConfig):
    names = ['norm_mha/gamma', 'norm_mha/beta', 'dense_qkv/kernel', 'dense_qkv/bias',
             'dense_mha/kernel', 'dense_mha/bias']
    ect = cfg.edge_channel_type
    if ect in ('residual', 'constrained'):
        names += ['norm_edge/gamma', 'norm_edge/beta']
    if ect != 'none':
        if cfg.gate_attention:
            names += ['attention_gates/kernel', 'attention_gates/bias']
        names += ['dense_edge_b/kernel', 'dense_edge_b/bias']
    if ect in ('residual', 'constrained'):
        names += ['dense_edge_r/kernel', 'dense_edge_r/bias']
    return names


def init_block_params(cfg: BlockConfig, seed: int = 1234, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Synthetic weights as SURVEY.md 8(d): Glorot-uniform kernels, biases U(-0.1,0.1),
    LN gamma U(0.5,1.5), beta U(-0.1,0.1)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in block_param_names(cfg):
        shape = PARAM_SHAPES[name](cfg)
        if name.endswith('kernel'):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            w = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        elif name.endswith('gamma'):
            w = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        else:
            w = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * 0.1
        out[name] = w.to(dtype)
    return out


def _dropout(x, rate, noise, training):
    if rate > 0 and training:
        assert noise is not None
        return x * (noise >= rate).to(x.dtype) / (1.0 - rate)
    return x


def mha_block(h, e, gates, mask, p, cfg: BlockConfig, edge_mask=None, training=False,
              noise: Optional[Dict[str, torch.Tensor]] = None):
    """graph_xformer_model_base.py:106-145.  ``e`` here is the per-head edge bias E
    (already projected) or None; returns (h', H_hat, A_tild)."""
    noise = noise or {}
    y = h                                                                   # :107
    if not cfg.add_n_norm:
        h = layer_norm(h, p['norm_mha/gamma'], p['norm_mha/beta'])          # :108-109
    qkv = dense(h, p['dense_qkv/kernel'], p['dense_qkv/bias'])              # :113-114
    edge_input = cfg.edge_channel_type != 'none'
    ins = [qkv] + ([e] if edge_input else []) + ([gates] if gates is not None else []) \
        + ([edge_mask] if edge_mask is not None else [])                    # :128-131
    h, e_out, mat = egt_layer(ins, mask=mask, training=training,
                              num_heads=cfg.num_heads,
                              clip_logits_value=cfg.clip_logits_value,
                              scale_degree=cfg.scale_degree,
                              edge_input=edge_input,
                              gate_input=gates is not None,
                              attn_mask=edge_mask is not None,
                              num_virtual_nodes=cfg.num_virtual_nodes,
                              random_mask_prob=cfg.random_mask_prob,
                              attn_dropout=cfg.attn_dropout,
                              scaler_type=cfg.scaler_type,
                              uniform_noise=noise.get('random_mask'),
                              dropout_noise=noise.get('attn_dropout'))      # :117-131
    h = dense(h, p['dense_mha/kernel'], p['dense_mha/bias'])                # :136-137
    h = _dropout(h, cfg.node_dropout, noise.get('node_dropout'), training)  # :138-139
    h = h + y                                                               # :140
    if cfg.add_n_norm:
        h = layer_norm(h, p['norm_mha/gamma'], p['norm_mha/beta'])          # :142-143
    return h, e_out, mat


def egt_block(h, e, mask, p, cfg: BlockConfig, edge_mask=None, training=False,
              noise: Optional[Dict[str, torch.Tensor]] = None, return_aux=False):
    """One attention block ``edge_update(tag, h, e) -> (h, e)``
    (graph_xformer_model_base.py:164-223, dispatch :328-339).

    h: [B,N,d]   e: [B,N,N,d_e]   mask: [B,N] bool (Keras mask of h)
    edge_mask: optional [B,N,N,h] 0/1 (``constrained`` variant)."""
    noise = noise or {}
    act = edge_activation_fn(cfg.edge_activation)
    ect = cfg.edge_channel_type
    if ect == 'none':                                                       # :164-171
        h2, H_hat, mat = mha_block(h, e, None, mask, p, cfg, edge_mask, training, noise)
        out = (h2, e)
    elif ect == 'bias':                                                     # :173-190
        e0 = e
        gates = None
        if cfg.gate_attention:
            gates = dense(e, p['attention_gates/kernel'], p['attention_gates/bias'])
        eb = act(dense(e, p['dense_edge_b/kernel'], p['dense_edge_b/bias']))
        h2, H_hat, mat = mha_block(h, eb, gates, mask, p, cfg, edge_mask, training, noise)
        out = (h2, e0)
    elif ect in ('residual', 'constrained'):                                # :192-223
        y = e
        if not cfg.add_n_norm:
            e = layer_norm(e, p['norm_edge/gamma'], p['norm_edge/beta'])    # :194-195
        gates = None
        if cfg.gate_attention:
            gates = dense(e, p['attention_gates/kernel'], p['attention_gates/bias'])   # :200-204
        eb = act(dense(e, p['dense_edge_b/kernel'], p['dense_edge_b/bias']))          # :208
        h2, H_hat, mat = mha_block(h, eb, gates, mask, p, cfg, edge_mask, training, noise)  # :212
        e2 = dense(H_hat, p['dense_edge_r/kernel'], p['dense_edge_r/bias'])           # :214-215
        e2 = _dropout(e2, cfg.edge_dropout, noise.get('edge_dropout'), training)      # :216-217
        e2 = e2 + y                                                                   # :218
        if cfg.add_n_norm:
            e2 = layer_norm(e2, p['norm_edge/gamma'], p['norm_edge/beta'])            # :220-221
        out = (h2, e2)
    else:
        raise KeyError(ect)
    if return_aux:
        return out + (dict(H_hat=H_hat, A_tild=mat),)
    return out


# --------------------------------------------------------------------------------------
# FFN half of a layer  ("next" row 8f-1;  graph_xformer_model_base.py:229-258,309-324)
# --------------------------------------------------------------------------------------
def _keras_activation(name):
    return dict(elu=torch.nn.functional.elu, relu=torch.relu, tanh=torch.tanh,
                sigmoid=torch.sigmoid, linear=lambda v: v)[name]


def ffn_channel(x, p, prefix, cfg: BlockConfig):
    """ffnlr1 + ffnlr2 for one channel (no cross-talk, dropout 0)."""
    y = x
    if not cfg.add_n_norm:
        x = layer_norm(x, p[f'{prefix}/norm/gamma'], p[f'{prefix}/norm/beta'])
    x = _keras_activation(cfg.activation)(dense(x, p[f'{prefix}/lr1/kernel'], p[f'{prefix}/lr1/bias']))
    x = dense(x, p[f'{prefix}/lr2/kernel'], p[f'{prefix}/lr2/bias'])
    x = x + y
    if cfg.add_n_norm:
        x = layer_norm(x, p[f'{prefix}/norm/gamma'], p[f'{prefix}/norm/beta'])
    return x


# --------------------------------------------------------------------------------------
# mask producers
# --------------------------------------------------------------------------------------
def node_mask_from_features(feat: torch.Tensor) -> torch.Tensor:
    """Neg1MaskedEmbedding.compute_mask (lib/base/xformer_layers/masking.py:35-43):
    valid <=> feature != -1."""
    return feat != -1


def prepend_virtual_nodes_mask(mask: torch.Tensor, n_vn: int) -> torch.Tensor:
    """VirtualNodeEmbedding.compute_mask (lib/base/graph_layers/virtual_nodes.py:47-50)."""
    return torch.cat([torch.ones(mask.shape[0], n_vn, dtype=torch.bool), mask], dim=1)


def edge_mask_from_adjacency(adj: torch.Tensor, num_heads: int, n_vn: int = 0) -> torch.Tensor:
    """AdjMatModel.get_edge_mask (lib/models/graph_model_base.py:131-142) then
    VNModel.get_edge_mask (:248-268): tile over heads, pad ones for virtual nodes."""
    m = adj[..., None].repeat(1, 1, 1, num_heads)
    if n_vn > 0:
        B, n1, n2, nh = m.shape
        row_true = torch.ones(B, n_vn, n2, nh, dtype=m.dtype)
        col_true = torch.ones(B, n1 + n_vn, n_vn, nh, dtype=m.dtype)
        m = torch.cat([row_true, m], dim=1)
        m = torch.cat([col_true, m], dim=2)
    return m


# --------------------------------------------------------------------------------------
# closed-form backward of the block -- the spec the CUDA backward follows.
# Checked against autograd in tests/test_oracle.py.
# --------------------------------------------------------------------------------------
def egt_block_backward(h, e, mask, p, cfg: BlockConfig, dh_out, de_out, edge_mask=None,
                       training=False, noise=None):
    """Hand-derived gradients of ``egt_block`` for edge_channel_type in
    {residual, constrained, bias, none}, pre-norm (add_n_norm=False), no node/edge dropout.

    Returns (dh, de, grads: dict name -> tensor)."""
    assert not cfg.add_n_norm
    noise = noise or {}
    ect = cfg.edge_channel_type
    B, N, d = h.shape
    nh = cfg.num_heads
    dk = d // nh
    scale = dk ** -0.5
    dt = h.dtype
    residual = ect in ('residual', 'constrained')
    edge_input = ect != 'none'
    gated = cfg.gate_attention and edge_input
    grads: Dict[str, torch.Tensor] = {}

    # ---------------- forward recompute (same order as the reference) ----------------
    def ln_fwd(x, g_, b_):
        mu = x.mean(-1, keepdim=True)
        xc = x - mu
        var = (xc * xc).mean(-1, keepdim=True)
        rstd = torch.rsqrt(var + LN_EPS)
        xhat = xc * rstd
        return xhat * g_ + b_, xhat, rstd

    def ln_bwd(dy, xhat, rstd, g_):
        dxhat = dy * g_
        dgamma = (dy * xhat).reshape(-1, dy.shape[-1]).sum(0)
        dbeta = dy.reshape(-1, dy.shape[-1]).sum(0)
        m1 = dxhat.mean(-1, keepdim=True)
        m2 = (dxhat * xhat).mean(-1, keepdim=True)
        dx = rstd * (dxhat - m1 - xhat * m2)
        return dx, dgamma, dbeta

    hn, h_xhat, h_rstd = ln_fwd(h, p['norm_mha/gamma'], p['norm_mha/beta'])
    qkv = dense(hn, p['dense_qkv/kernel'], p['dense_qkv/bias'])
    Q, K, V = qkv.reshape(B, N, 3, dk, nh).unbind(2)

    if residual:
        en, e_xhat, e_rstd = ln_fwd(e, p['norm_edge/gamma'], p['norm_edge/beta'])
    else:
        en = e
    if edge_input:
        Epre = dense(en, p['dense_edge_b/kernel'], p['dense_edge_b/bias'])
        act = edge_activation_fn(cfg.edge_activation)
        Eb = act(Epre)
    if gated:
        G = dense(en, p['attention_gates/kernel'], p['attention_gates/bias'])

    S = torch.einsum('bldh,bmdh->blmh', Q, K) * scale
    if cfg.clip_logits_value is not None:
        lo, hi = cfg.clip_logits_value
        inside = ((S >= lo) & (S <= hi)).to(dt)
        A_hat = torch.clamp(S, lo, hi)
    else:
        inside = torch.ones_like(S)
        A_hat = S
    H_hat = A_hat + Eb if edge_input else A_hat

    neg = torch.zeros(B, N, N, nh, dtype=dt)
    if mask is not None:
        neg = neg + (mask[:, None, :, None].to(dt) - 1) * NEG_INF_MASK
    if edge_mask is not None:
        neg = neg + (edge_mask.to(dt) - 1) * NEG_INF_MASK
    if cfg.random_mask_prob > 0 and training:
        neg = neg + torch.where(noise['random_mask'].to(dt) < cfg.random_mask_prob,
                                torch.full_like(neg, -NEG_INF_MASK), torch.zeros_like(neg))
    P = torch.softmax(H_hat + neg, dim=2)
    if gated:
        g = torch.sigmoid(G + neg)
        A = P * g
    else:
        g = None
        A = P
    keepscale = None
    if cfg.attn_dropout > 0 and training:
        keepscale = (noise['attn_dropout'] >= cfg.attn_dropout).to(dt) / (1 - cfg.attn_dropout)
        A_d = A * keepscale
    else:
        A_d = A
    O = torch.einsum('blmh,bmdh->bldh', A_d, V)
    if cfg.scale_degree:
        deg = g.sum(2)                                       # b,l,h
        s = torch.log(1 + deg) if cfg.scaler_type == 'log' else deg
        if cfg.num_virtual_nodes > 0:
            s = torch.cat([torch.ones_like(s[:, :cfg.num_virtual_nodes]),
                           s[:, cfg.num_virtual_nodes:]], dim=1)
        V_att = O * s[:, :, None, :]
    else:
        V_att = O
    V_att_flat = V_att.reshape(B, N, d)

    # ---------------- backward ----------------
    # h' = V_att W_O + b_O + h
    grads['dense_mha/kernel'] = V_att_flat.reshape(-1, d).T @ dh_out.reshape(-1, d)
    grads['dense_mha/bias'] = dh_out.reshape(-1, d).sum(0)
    dV_att = (dh_out @ p['dense_mha/kernel'].T).reshape(B, N, dk, nh)

    if cfg.scale_degree:
        dO = dV_att * s[:, :, None, :]
        ds = (dV_att * O).sum(2)                             # b,l,h
        if cfg.scaler_type == 'log':
            ddeg = ds / (1 + deg)
        else:
            ddeg = ds.clone()
        if cfg.num_virtual_nodes > 0:
            ddeg[:, :cfg.num_virtual_nodes] = 0
    else:
        dO = dV_att
        ddeg = None

    dA_d = torch.einsum('bldh,bmdh->blmh', dO, V)
    dV = torch.einsum('blmh,bldh->bmdh', A_d, dO)
    dA = dA_d * keepscale if keepscale is not None else dA_d
    if gated:
        dP = dA * g
        dg = dA * P
        if ddeg is not None:
            dg = dg + ddeg[:, :, None, :]
        dG = dg * g * (1 - g)
    else:
        dP = dA
        dG = None
    dHm = P * (dP - (dP * P).sum(2, keepdim=True))           # softmax backward

    if residual:
        # e' = H_hat W_r + b_r + e
        Wr = p['dense_edge_r/kernel']
        grads['dense_edge_r/kernel'] = H_hat.reshape(-1, nh).T @ de_out.reshape(-1, cfg.edge_width)
        grads['dense_edge_r/bias'] = de_out.reshape(-1, cfg.edge_width).sum(0)
        dH = dHm + de_out @ Wr.T
    else:
        dH = dHm
    dS = dH * inside
    dQ = torch.einsum('blmh,bmdh->bldh', dS, K) * scale
    dK = torch.einsum('blmh,bldh->bmdh', dS, Q) * scale

    de = torch.zeros_like(e) if e is not None else None
    if edge_input:
        dE = dH
        # through the activation
        Epre_ = Epre.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            act_out = edge_activation_fn(cfg.edge_activation)(Epre_)
            (dEpre,) = torch.autograd.grad(act_out, Epre_, dE)
        WE = p['dense_edge_b/kernel']
        grads['dense_edge_b/kernel'] = en.reshape(-1, cfg.edge_width).T @ dEpre.reshape(-1, nh)
        grads['dense_edge_b/bias'] = dEpre.reshape(-1, nh).sum(0)
        den = dEpre @ WE.T
        if gated:
            WG = p['attention_gates/kernel']
            grads['attention_gates/kernel'] = en.reshape(-1, cfg.edge_width).T @ dG.reshape(-1, nh)
            grads['attention_gates/bias'] = dG.reshape(-1, nh).sum(0)
            den = den + dG @ WG.T
        if residual:
            de_ln, dgam, dbet = ln_bwd(den, e_xhat, e_rstd, p['norm_edge/gamma'])
            grads['norm_edge/gamma'] = dgam
            grads['norm_edge/beta'] = dbet
            de = de_ln + de_out
        else:                                               # 'bias': e0 is returned unchanged
            de = den + de_out
    else:
        de = de_out.clone() if de_out is not None else None

    dqkv = torch.stack([dQ, dK, dV], dim=2).reshape(B, N, 3 * d)
    grads['dense_qkv/kernel'] = hn.reshape(-1, d).T @ dqkv.reshape(-1, 3 * d)
    grads['dense_qkv/bias'] = dqkv.reshape(-1, 3 * d).sum(0)
    dhn = dqkv @ p['dense_qkv/kernel'].T
    dh_ln, dgam, dbet = ln_bwd(dhn, h_xhat, h_rstd, p['norm_mha/gamma'])
    grads['norm_mha/gamma'] = dgam
    grads['norm_mha/beta'] = dbet
    dh = dh_ln + dh_out
    return dh, de, grads


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------
def synthetic_batch(B, N, d, d_e, seed=20240, ragged=False, dtype=torch.float32, min_frac=0.5):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(B, N, d, generator=g, dtype=torch.float32).to(dtype)
    e = torch.randn(B, N, N, d_e, generator=g, dtype=torch.float32).to(dtype)
    if ragged:
        lo = max(1, int(math.ceil(N * min_frac)))
        nn_ = torch.randint(lo, N + 1, (B,), generator=g)
    else:
        nn_ = torch.full((B,), N, dtype=torch.int64)
    mask = torch.arange(N)[None, :] < nn_[:, None]
    return h, e, mask
