"""torch.autograd bindings over the C ABI.  PyTorch is plumbing here: device memory, streams,
autograd bookkeeping.  All arithmetic happens in libegt_b200.so."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import _lib as L


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.EGT_F32
    if t.dtype == torch.bfloat16:
        return L.EGT_BF16
    raise TypeError(f'egt_b200 supports float32 and bfloat16 activations, got {t.dtype}')


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('egt_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback')


def _refuse_rng_capture(training, random_mask_prob, attn_dropout, offset_dev=None):
    """The Philox (seed, offset) of the random key mask / attention dropout reach the kernels as launch ARGUMENTS.
    A CUDA graph freezes its launch arguments, so every replay of a captured training step would draw the identical
    mask -- silently different from the reference, which draws fresh noise per step (egt_layers.py:103-108,116-117).
    With ``offset_dev`` (a device word the kernels add to the offset when they run; the layers of egt_b200.layers keep
    one per module and bump it inside the captured step) a replay draws new noise and capture is fine.  Without it
    capture is refused while that RNG is live; EGT_ALLOW_FROZEN_RNG=1 accepts a frozen mask (e.g. to time the kernels)."""
    import os
    if training and (random_mask_prob > 0 or attn_dropout > 0) and offset_dev is None \
            and torch.cuda.is_current_stream_capturing() and os.environ.get('EGT_ALLOW_FROZEN_RNG', '0') != '1':
        raise RuntimeError('egt_b200: refusing to capture a training step with random_mask_prob / attn_dropout > 0 into a '
                           'CUDA graph without a device-side RNG offset: every replay would reuse the same mask')


def graph_safe_offset(counter: torch.Tensor):
    """For a step that is being captured into a CUDA graph: returns a snapshot of the module's device RNG counter (an
    int64 tensor of one element that must exist BEFORE the capture) and bumps the counter -- both as captured work, so
    every replay sees the next value.  The snapshot is what the forward AND the backward kernels of this call add to
    their Philox offset."""
    snap = counter.clone()
    counter.add_(1)
    return snap


def _mask_u8(mask, B, N):
    if mask is None:
        return None
    if isinstance(mask, (list, tuple)):          # Keras passes one mask per input; only mask[0] is used
        mask = mask[0]                           # (egt_layers.py:66)
    if mask is None:
        return None
    assert mask.shape == (B, N), f'mask must be [B,N]={B, N}, got {tuple(mask.shape)}'
    m = mask.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)


# ------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class AttnSpec:
    """Constructor arguments of the reference ``EGT`` layer (egt_layers.py:5-16)."""
    num_heads: int = 8
    clip_logits_value: Optional[Sequence[float]] = (-5., 5.)
    scale_degree: bool = False
    scaler_type: str = 'log'
    edge_input: bool = True
    gate_input: bool = True
    attn_mask: bool = False
    num_virtual_nodes: int = 0
    random_mask_prob: float = 0.
    attn_dropout: float = 0.

    def validate(self):
        if self.scale_degree and not self.gate_input:                       # egt_layers.py:20-21
            raise ValueError('scale_degree requires gate_input')
        if self.scaler_type not in ('log', 'linear'):                       # :23-24
            raise ValueError('scaler_type must be log or linear')

    def c_cfg(self, B, N, dk, dtype, training, seed, offset, mask_kind=None, offset_dev=None) -> L.AttnCfg:
        c = L.AttnCfg()
        c.B, c.N, c.h, c.dk, c.dtype = B, N, self.num_heads, dk, dtype
        c.edge_input, c.gate_input = int(self.edge_input), int(self.gate_input)
        c.attn_mask = (L.EGT_MASK_DENSE if mask_kind is None else mask_kind) if self.attn_mask else L.EGT_MASK_NONE
        c.has_clip = int(self.clip_logits_value is not None)
        if self.clip_logits_value is not None:
            c.clip_lo, c.clip_hi = float(self.clip_logits_value[0]), float(self.clip_logits_value[1])
        c.scale_degree = int(self.scale_degree)
        c.scaler_type = L.EGT_SCALER_LOG if self.scaler_type == 'log' else L.EGT_SCALER_LINEAR
        c.num_virtual_nodes = self.num_virtual_nodes
        c.training = int(bool(training))
        c.random_mask_prob, c.attn_dropout = float(self.random_mask_prob), float(self.attn_dropout)
        c.seed, c.offset = int(seed) & (2**64 - 1), int(offset) & (2**64 - 1)
        c.offset_dev = None if offset_dev is None else offset_dev.data_ptr()
        return c


class _EGTAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, E, G, M, mask_u8, spec: AttnSpec, training, seed, offset, want_attn, offset_dev=None):
        lib = L.load()
        _need_cuda(qkv, E, G, M, mask_u8)
        B, N, C3 = qkv.shape
        h = spec.num_heads
        assert C3 % (3 * h) == 0, 'qkv channels must be divisible by 3*num_heads'    # egt_layers.py:70
        dk = C3 // (3 * h)
        d = dk * h
        qkv = qkv.contiguous()
        E = None if E is None else E.contiguous()
        G = None if G is None else G.contiguous()
        M = None if M is None else M.to(qkv.dtype).contiguous()
        for t, nm in ((E, 'E'), (G, 'G'), (M, 'M')):
            if t is not None:
                assert t.shape == (B, N, N, h) and t.dtype == qkv.dtype, f'{nm} must be [B,N,N,h] {qkv.dtype}'
        cfg = spec.c_cfg(B, N, dk, _dtype_code(qkv), training, seed, offset, offset_dev=offset_dev)
        ctx.offset_dev = offset_dev                      # keeps the snapshot alive until the backward has used it
        v_att = torch.empty(B, N, d, dtype=qkv.dtype, device=qkv.device)
        h_hat = torch.empty(B, N, N, h, dtype=qkv.dtype, device=qkv.device)
        a_tild = torch.empty(B, N, N, h, dtype=qkv.dtype, device=qkv.device) if want_attn else None
        lse = torch.empty(2, B, N, h, dtype=torch.float32, device=qkv.device)
        deg = torch.empty(B, N, h, dtype=torch.float32, device=qkv.device)
        L.check(lib.egt_attn_fwd(C.byref(cfg), _ptr(qkv), _ptr(E), _ptr(G), _ptr(M), _ptr(mask_u8),
                                 _ptr(v_att), _ptr(h_hat), _ptr(a_tild), _ptr(lse), _ptr(deg), _stream()))
        ctx.save_for_backward(qkv, E, G, M, mask_u8, lse, deg)
        ctx.cfg = cfg
        if want_attn:
            ctx.mark_non_differentiable(a_tild)
            return v_att, h_hat, a_tild
        return v_att, h_hat

    @staticmethod
    def backward(ctx, d_v_att, d_h_hat, *unused):
        lib = L.load()
        qkv, E, G, M, mask_u8, lse, deg = ctx.saved_tensors
        cfg = ctx.cfg
        B, N, h = cfg.B, cfg.N, cfg.h
        if d_v_att is None:
            d_v_att = torch.zeros(B, N, h * cfg.dk, dtype=qkv.dtype, device=qkv.device)
        d_v_att = d_v_att.contiguous()
        d_h_hat = None if d_h_hat is None else d_h_hat.contiguous()
        d_qkv = torch.empty_like(qkv)
        dE = torch.empty_like(E) if E is not None else None
        dG = torch.empty_like(G) if G is not None else None
        row_ws = torch.empty(2, B, N, h, dtype=torch.float32, device=qkv.device)
        L.check(lib.egt_attn_bwd(C.byref(cfg), _ptr(qkv), _ptr(E), _ptr(G), _ptr(M), _ptr(mask_u8),
                                 _ptr(lse), _ptr(deg), _ptr(d_v_att), _ptr(d_h_hat), _ptr(d_qkv),
                                 _ptr(dE), _ptr(dG), _ptr(row_ws), _stream()))
        return d_qkv, dE, dG, None, None, None, None, None, None, None, None


def egt_attention(inputs, mask=None, training=False, *, spec: AttnSpec, seed=0, offset=0, return_attn=False, offset_dev=None):
    """``EGT.call``: ([QKV, E?, G?, M?], mask, training) -> (V_att, H_hat, A_tild | None)
    (egt_layers.py:57-143 / :145-213).  Positional input order as in egt_layers.py:62-65."""
    spec.validate()
    inputs = list(inputs)
    qkv = inputs.pop(0)
    E = inputs.pop(0) if spec.edge_input else None
    G = inputs.pop(0) if spec.gate_input else None
    M = inputs.pop(0) if spec.attn_mask else None
    B, N, _ = qkv.shape
    m8 = _mask_u8(mask, B, N)
    _refuse_rng_capture(training, spec.random_mask_prob, spec.attn_dropout, offset_dev)
    out = _EGTAttnFn.apply(qkv, E, G, M, m8, spec, training, seed, offset, return_attn, offset_dev)
    if return_attn:
        return out
    return out[0], out[1], None


# ------------------------------------------------------------------------------------------
_ACT = {None: L.EGT_ACT_NONE, 'linear': L.EGT_ACT_NONE, 'relu': L.EGT_ACT_RELU, 'elu': L.EGT_ACT_ELU,
        'tanh': L.EGT_ACT_TANH, 'sigmoid': L.EGT_ACT_SIGMOID}
_ECT = {'none': L.EGT_EDGE_NONE, 'bias': L.EGT_EDGE_BIAS, 'residual': L.EGT_EDGE_RESIDUAL,
        'constrained': L.EGT_EDGE_CONSTRAINED}


@dataclass(frozen=True)
class BlockSpec:
    """The GraphTransformerBase constructor arguments that reach the attention block
    (graph_xformer_model_base.py:17-45), same names and defaults."""
    model_width: int = 128
    edge_width: int = 32
    num_heads: int = 8
    gate_attention: bool = True
    node_dropout: float = 0.
    edge_dropout: float = 0.
    add_n_norm: bool = False
    clip_logits_value: Optional[Sequence[float]] = (-5., 5.)
    edge_activation: Optional[str] = None
    edge_channel_type: str = 'residual'
    scale_degree: bool = False
    scaler_type: str = 'log'
    num_virtual_nodes: int = 0
    random_mask_prob: float = 0.
    attn_dropout: float = 0.
    ln_eps: float = 1e-3            # keras LayerNormalization default

    def validate(self):
        if not self.gate_attention and self.scale_degree:                   # graph_xformer_model_base.py:46-47
            raise ValueError('scale_degree only works with gate_attention')
        if self.edge_channel_type not in _ECT:
            raise KeyError(self.edge_channel_type)                          # dispatch dict :328-334
        if self.model_width % self.num_heads:
            raise AssertionError('model_width must be divisible by num_heads')
        if self.add_n_norm:
            raise NotImplementedError('add_n_norm=True (post-norm) is not built: no reference config uses it')
        if self.node_dropout or self.edge_dropout:
            raise NotImplementedError('node/edge dropout > 0 is not built: every reference config uses 0')

    @property
    def has_edge(self):
        return self.edge_channel_type != 'none'

    @property
    def is_residual(self):
        return self.edge_channel_type in ('residual', 'constrained')

    def act_code(self):
        a = self.edge_activation
        if a is not None and a.lower().startswith('lrelu'):                 # graph_xformer_model_base.py:150-152
            return L.EGT_ACT_LRELU, float(a[-1]) / 10
        key = a.lower() if isinstance(a, str) else a
        if key not in _ACT:
            raise ValueError(f'unsupported edge_activation {a!r}')
        return _ACT[key], 0.

    def c_cfg(self, B, N, dtype, training, seed, offset, offset_dev=None) -> L.BlockCfg:
        c = L.BlockCfg()
        a = c.attn
        a.B, a.N, a.h, a.dk, a.dtype = B, N, self.num_heads, self.model_width // self.num_heads, dtype
        a.has_clip = int(self.clip_logits_value is not None)
        if self.clip_logits_value is not None:
            a.clip_lo, a.clip_hi = float(self.clip_logits_value[0]), float(self.clip_logits_value[1])
        a.scale_degree = int(self.scale_degree)
        a.scaler_type = L.EGT_SCALER_LOG if self.scaler_type == 'log' else L.EGT_SCALER_LINEAR
        a.num_virtual_nodes = self.num_virtual_nodes
        a.training = int(bool(training))
        a.random_mask_prob, a.attn_dropout = float(self.random_mask_prob), float(self.attn_dropout)
        a.seed, a.offset = int(seed) & (2**64 - 1), int(offset) & (2**64 - 1)
        a.offset_dev = None if offset_dev is None else offset_dev.data_ptr()
        c.d_e = self.edge_width
        c.edge_channel_type = _ECT[self.edge_channel_type]
        c.gate_attention = int(self.gate_attention)
        c.edge_act, c.edge_act_alpha = self.act_code()
        c.ln_eps = self.ln_eps
        return c


def param_layout(spec: BlockSpec):
    """(total floats, {field: (offset, shape)}) of the flat parameter / gradient buffer."""
    lib = L.load()
    cfg = spec.c_cfg(1, 1, L.EGT_F32, False, 0, 0)
    offs = (C.c_int64 * 14)()
    total = lib.egt_block_param_layout(C.byref(cfg), offs)
    d, de, h = spec.model_width, spec.edge_width, spec.num_heads
    shapes = [(d,), (d,), (d, 3 * d), (3 * d,), (d, d), (d,), (de,), (de,), (de, h), (h,), (de, h), (h,),
              (h, de), (de,)]
    out = {}
    for name, off, shp in zip(L.WEIGHT_FIELDS, offs, shapes):
        if off >= 0:
            out[name] = (int(off), shp)
    return int(total), out


def _fill_ptrs(struct, flat: torch.Tensor, layout):
    base = flat.data_ptr()
    for name in L.WEIGHT_FIELDS:
        setattr(struct, name, base + 4 * layout[name][0] if name in layout else None)
    return struct


class _EGTBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, e, flat, mask_u8, adj_u8, spec: BlockSpec, layout, training, seed, offset, offset_dev=None):
        lib = L.load()
        _need_cuda(h, e, flat, mask_u8, adj_u8)
        B, N, d = h.shape
        assert d == spec.model_width, f'h has {d} channels, spec.model_width={spec.model_width}'
        assert flat.dtype == torch.float32 and flat.is_contiguous()
        h = h.contiguous()
        if spec.has_edge:
            assert e is not None and e.shape == (B, N, N, spec.edge_width) and e.dtype == h.dtype
            e = e.contiguous()
        cfg = spec.c_cfg(B, N, _dtype_code(h), training, seed, offset, offset_dev)
        ctx.offset_dev = offset_dev                      # keeps the snapshot alive until the backward has used it
        dev = h.device
        io = L.BlockFwdIO()
        h_out = torch.empty_like(h)
        e_out = torch.empty_like(e) if spec.is_residual else None
        qkv = torch.empty(B, N, 3 * d, dtype=h.dtype, device=dev)
        v_att = torch.empty(B, N, d, dtype=h.dtype, device=dev)
        lse = torch.empty(2, B, N, spec.num_heads, dtype=torch.float32, device=dev)
        deg = torch.empty(B, N, spec.num_heads, dtype=torch.float32, device=dev)
        nbytes = lib.egt_block_workspace_bytes(C.byref(cfg), 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        io.h, io.e, io.mask, io.adj = h.data_ptr(), (e.data_ptr() if spec.has_edge else None), \
            (mask_u8.data_ptr() if mask_u8 is not None else None), (adj_u8.data_ptr() if adj_u8 is not None else None)
        io.h_out, io.e_out = h_out.data_ptr(), (e_out.data_ptr() if e_out is not None else None)
        io.qkv, io.v_att, io.lse, io.deg = qkv.data_ptr(), v_att.data_ptr(), lse.data_ptr(), deg.data_ptr()
        io.workspace, io.workspace_bytes = ws.data_ptr(), nbytes
        w = _fill_ptrs(L.BlockWeights(), flat, layout)
        L.check(lib.egt_block_fwd(C.byref(cfg), C.byref(w), C.byref(io), _stream()))
        ctx.save_for_backward(h, e if spec.has_edge else None, flat, mask_u8, adj_u8, qkv, v_att, lse, deg)
        ctx.cfg, ctx.spec, ctx.layout = cfg, spec, layout
        ctx.path = lib.egt_last_path()
        if spec.is_residual:
            return h_out, e_out
        return h_out

    @staticmethod
    def backward(ctx, dh_out, de_out=None):
        lib = L.load()
        h, e, flat, mask_u8, adj_u8, qkv, v_att, lse, deg = ctx.saved_tensors
        spec, cfg, layout = ctx.spec, ctx.cfg, ctx.layout
        dev = h.device
        dh_out = dh_out.contiguous()
        de_out = None if de_out is None else de_out.contiguous()
        dh = torch.empty_like(h)
        de = torch.empty_like(e) if spec.has_edge else None
        dflat = torch.zeros_like(flat)
        nbytes = lib.egt_block_workspace_bytes(C.byref(cfg), 1)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        io = L.BlockBwdIO()
        io.h, io.e = h.data_ptr(), (e.data_ptr() if e is not None else None)
        io.mask = mask_u8.data_ptr() if mask_u8 is not None else None
        io.adj = adj_u8.data_ptr() if adj_u8 is not None else None
        io.qkv, io.v_att, io.lse, io.deg = qkv.data_ptr(), v_att.data_ptr(), lse.data_ptr(), deg.data_ptr()
        io.dh_out = dh_out.data_ptr()
        io.de_out = de_out.data_ptr() if de_out is not None else None
        io.dh, io.de = dh.data_ptr(), (de.data_ptr() if de is not None else None)
        io.workspace, io.workspace_bytes = ws.data_ptr(), nbytes
        w = _fill_ptrs(L.BlockWeights(), flat, layout)
        g = _fill_ptrs(L.BlockGrads(), dflat, layout)
        L.check(lib.egt_block_bwd(C.byref(cfg), C.byref(w), C.byref(g), C.byref(io), _stream()))
        return dh, de, dflat, None, None, None, None, None, None, None, None


def egt_block(h, e, mask, flat, spec: BlockSpec, layout, edge_mask=None, training=False, seed=0, offset=0, offset_dev=None):
    """``edge_update(tag, h, e) -> (h, e)`` (graph_xformer_model_base.py:164-223, :328-339)."""
    spec.validate()
    B, N, _ = h.shape
    m8 = _mask_u8(mask, B, N)
    adj = None
    if spec.edge_channel_type == 'constrained':
        if edge_mask is None:
            raise ValueError("edge_channel_type='constrained' needs edge_mask")
        if edge_mask.dim() == 4:      # reference form [B,N,N,h], tiled over heads (graph_model_base.py:139-141)
            edge_mask = edge_mask[..., 0]
        adj = (edge_mask != 0).to(torch.uint8).contiguous()
    e_in = e if spec.has_edge else None
    _refuse_rng_capture(training, spec.random_mask_prob, spec.attn_dropout, offset_dev)
    out = _EGTBlockFn.apply(h, e_in, flat, m8, adj, spec, layout, training, seed, offset, offset_dev)
    if spec.is_residual:
        return out
    return out, e


# ------------------------------------------------------------------------------------------
# feed-forward half of a layer ("next" row)
# ------------------------------------------------------------------------------------------
def ffn_layout(width: int, hidden: int):
    """(total floats, {field: (offset, shape)}) of one channel's flat FFN parameter buffer."""
    shapes = [(width,), (width,), (width, hidden), (hidden,), (hidden, width), (width,)]
    out, off = {}, 0
    for name, shp in zip(L.FFN_FIELDS, shapes):
        n = 1
        for v in shp:
            n *= v
        out[name] = (off, shp)
        off += (n + 3) // 4 * 4                      # keep every tensor 16-byte aligned
    return off, out


class _EGTFfnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flat, width, hidden, act_code, ln_eps, layout):
        lib = L.load()
        _need_cuda(x, flat)
        assert x.shape[-1] == width, f'x has {x.shape[-1]} channels, width={width}'
        x = x.contiguous()
        cfg = L.FfnCfg()
        cfg.rows, cfg.width, cfg.hidden = x.numel() // width, width, hidden
        cfg.dtype, cfg.activation, cfg.ln_eps = _dtype_code(x), act_code, ln_eps
        y = torch.empty_like(x)
        w = L.FfnWeights()
        for f in L.FFN_FIELDS:
            setattr(w, f, flat.data_ptr() + 4 * layout[f][0])
        nws = int(lib.egt_ffn_workspace_bytes(C.byref(cfg)))
        ws = torch.empty(nws, dtype=torch.uint8, device=x.device) if nws else None
        L.check(lib.egt_ffn_fwd_ws(C.byref(cfg), C.byref(w), _ptr(x), _ptr(y), _ptr(ws) if nws else None, nws, _stream()))
        ctx.save_for_backward(x, flat)
        ctx.cfg, ctx.layout = cfg, layout
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, flat = ctx.saved_tensors
        cfg, layout = ctx.cfg, ctx.layout
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dflat = torch.zeros_like(flat)
        w, g = L.FfnWeights(), L.FfnGrads()
        for f in L.FFN_FIELDS:
            setattr(w, f, flat.data_ptr() + 4 * layout[f][0])
            setattr(g, f, dflat.data_ptr() + 4 * layout[f][0])
        nws = int(lib.egt_ffn_workspace_bytes(C.byref(cfg)))
        ws = torch.empty(nws, dtype=torch.uint8, device=x.device) if nws else None
        L.check(lib.egt_ffn_bwd_ws(C.byref(cfg), C.byref(w), C.byref(g), _ptr(x), _ptr(dy), _ptr(dx), _ptr(ws) if nws else None, nws,
                                   _stream()))
        return dx, dflat, None, None, None, None, None


def egt_ffn(x, flat, width, hidden, layout, activation='elu', ln_eps=1e-3):
    """``ffnlr1 -> ffnact -> ffnlr2`` for one channel (graph_xformer_model_base.py:229-258)."""
    key = activation.lower() if isinstance(activation, str) else activation
    if key not in _ACT:
        raise ValueError(f'unsupported activation {activation!r}')
    return _EGTFfnFn.apply(x, flat, width, hidden, _ACT[key], ln_eps, layout)
