"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE -- test infrastructure.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The reference files are imported unmodified:

  /root/reference/lib/models/egt_layers.py                (class EGT)
  /root/reference/lib/models/graph_xformer_model_base.py  (GraphTransformerBase.transform_embeddings:
                                                           mha_block, edge_update_*, ffn_block)

on top of the TF/Keras shim in oracle/tf_shim (TensorFlow itself is not installable here).
For every case the restatement in oracle/egt_oracle.py is checked against the reference
output before the vectors are written, so a committed fixture is by construction one that
both agree on.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
"""
import itertools
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('EGT_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'tf_shim'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import tensorflow as tf                                   # the shim  # noqa: E402
from tensorflow import keras                              # noqa: E402
from lib.models.egt_layers import EGT                     # reference source  # noqa: E402
from lib.models.graph_xformer_model_base import GraphTransformerBase  # noqa: E402

from oracle import egt_oracle as O                        # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _np(x):
    return x.detach().cpu().numpy()


def layer_cases():
    """EGT-layer cases: every flag the layer has (egt_layers.py:5-16)."""
    cases = []
    base = dict(B=2, N=7, h=4, dk=3)
    for gated, edge_in, attn_m in itertools.product([True, False], [True, False], [True, False]):
        cases.append(dict(base, gate_input=gated, edge_input=edge_in, attn_mask=attn_m))
    cases.append(dict(base, gate_input=True, edge_input=True, attn_mask=False, scale_degree=True, scaler_type='log'))
    cases.append(dict(base, gate_input=True, edge_input=True, attn_mask=False, scale_degree=True, scaler_type='linear'))
    cases.append(dict(base, gate_input=True, edge_input=True, attn_mask=True, scale_degree=True, scaler_type='log',
                      num_virtual_nodes=2))
    cases.append(dict(base, gate_input=True, edge_input=True, attn_mask=False, clip_logits_value=None))
    cases.append(dict(base, gate_input=True, edge_input=True, attn_mask=False, random_mask_prob=0.3, training=1))
    cases.append(dict(base, gate_input=False, edge_input=True, attn_mask=False, random_mask_prob=0.3, training=1))
    cases.append(dict(base, gate_input=True, edge_input=True, attn_mask=False, attn_dropout=0.25, training=1))
    cases.append(dict(base, gate_input=True, edge_input=True, attn_mask=False, random_mask_prob=0.3, training=0))
    cases.append(dict(B=3, N=12, h=8, dk=8, gate_input=True, edge_input=True, attn_mask=False,
                      scale_degree=True, random_mask_prob=0.1, training=1))
    # the all-keys-masked corner: p=0.97 random masking on tiny graphs
    cases.append(dict(B=2, N=3, h=2, dk=2, gate_input=False, edge_input=True, attn_mask=False,
                      random_mask_prob=0.97, training=1))
    cases.append(dict(B=2, N=3, h=2, dk=2, gate_input=True, edge_input=True, attn_mask=False,
                      random_mask_prob=0.97, training=1, scale_degree=True))
    return cases


def run_layer_case(idx, c, dtype=torch.float64):
    c = dict(c)
    B, N, h, dk = c.pop('B'), c.pop('N'), c.pop('h'), c.pop('dk')
    training = c.pop('training', 0)
    g = torch.Generator().manual_seed(100 + idx)
    QKV = torch.randn(B, N, 3 * dk * h, generator=g, dtype=dtype) * 1.5
    E = torch.randn(B, N, N, h, generator=g, dtype=dtype)
    G = torch.randn(B, N, N, h, generator=g, dtype=dtype)
    adj = (torch.rand(B, N, N, generator=g) < 0.5) | torch.eye(N, dtype=torch.bool)[None]
    M = adj[..., None].repeat(1, 1, 1, h).to(dtype)
    nn_ = torch.randint(max(1, N // 2), N + 1, (B,), generator=g)
    mask = torch.arange(N)[None] < nn_[:, None]

    flags = dict(num_heads=h, clip_logits_value=[-5., 5.], scale_degree=False, scaler_type='log',
                 edge_input=True, gate_input=True, attn_mask=False, num_virtual_nodes=0,
                 random_mask_prob=0.0, attn_dropout=0.0)
    flags.update(c)
    ins = [QKV.clone()]
    ins[0]._keras_mask = mask
    if flags['edge_input']:
        ins.append(E)
    if flags['gate_input']:
        ins.append(G)
    if flags['attn_mask']:
        ins.append(M)

    keras.backend.set_learning_phase(training)
    tf.random.seed(7000 + idx)
    layer = EGT(name=f'mha_case{idx}', **flags)
    V_att, H_hat, A_tild = layer(ins)
    draws = dict()
    for tag, u in tf.random.draws:
        draws[tag] = u
    keras.backend.set_learning_phase(0)

    # restatement must agree with the reference source
    o_V, o_H, o_A = O.egt_layer(ins, mask=mask, training=bool(training),
                                uniform_noise=draws.get('uniform'), dropout_noise=draws.get('dropout'),
                                **{k: (tuple(v) if isinstance(v, list) else v) for k, v in flags.items()})
    for a, b, nm in ((V_att, o_V, 'V_att'), (H_hat, o_H, 'H_hat'), (A_tild, o_A, 'A_tild')):
        err = (a - b).abs().max().item()
        assert err < 1e-12, f'layer case {idx} {nm}: restatement differs from reference source by {err}'

    rec = dict(QKV=_np(QKV), E=_np(E), G=_np(G), M=_np(M), mask=_np(mask),
               V_att=_np(V_att), H_hat=_np(H_hat), A_tild=_np(A_tild), training=np.int64(training))
    for k, v in flags.items():
        if k == 'clip_logits_value':
            rec['flag_clip'] = np.array([np.nan, np.nan] if v is None else v, dtype=np.float64)
        elif k == 'scaler_type':
            rec['flag_scaler_type'] = np.array(v)
        else:
            rec['flag_' + k] = np.array(v)
    if 'uniform' in draws:
        rec['uniform_noise'] = _np(draws['uniform'])
    if 'dropout' in draws:
        rec['dropout_noise'] = _np(draws['dropout'])
    return rec


def block_cases():
    cases = []
    base = dict(B=2, N=9, model_width=16, edge_width=8, num_heads=4)
    cases.append(dict(base))
    cases.append(dict(base, edge_channel_type='bias'))
    cases.append(dict(base, edge_channel_type='none'))
    cases.append(dict(base, edge_channel_type='constrained'))
    cases.append(dict(base, gate_attention=False))
    cases.append(dict(base, scale_degree=True))
    cases.append(dict(base, scale_degree=True, scaler_type='linear', num_virtual_nodes=1))
    cases.append(dict(base, edge_activation='lrelu2'))
    cases.append(dict(base, edge_activation='elu'))
    cases.append(dict(base, add_n_norm=True))
    cases.append(dict(base, random_mask_prob=0.2, training=1))
    cases.append(dict(base, clip_logits_value=None))
    cases.append(dict(B=2, N=12, model_width=64, edge_width=8, num_heads=8, scale_degree=True))
    cases.append(dict(B=1, N=13, model_width=48, edge_width=48, num_heads=8))          # ZINC 100K widths, dk=6
    cases.append(dict(B=1, N=10, model_width=96, edge_width=8, num_heads=8))           # dk=12
    cases.append(dict(B=1, N=11, model_width=128, edge_width=32, num_heads=16, scale_degree=True))
    return cases


def run_block_case(idx, c, dtype=torch.float64):
    c = dict(c)
    B, N = c.pop('B'), c.pop('N')
    training = c.pop('training', 0)
    keras._STATE['seed'] = 4321 + idx
    keras._STATE['counter'] = 0
    model = GraphTransformerBase(model_height=1, **c)      # reference constructor + defaults
    cfg = model.config
    d, de, nh = cfg.model_width, cfg.edge_width, cfg.num_heads
    g = torch.Generator().manual_seed(500 + idx)
    h = torch.randn(B, N, d, generator=g, dtype=dtype)
    e = torch.randn(B, N, N, de, generator=g, dtype=dtype)
    nn_ = torch.randint(max(1, N // 2), N + 1, (B,), generator=g)
    mask = torch.arange(N)[None] < nn_[:, None]
    edge_mask = None
    if cfg.edge_channel_type == 'constrained':
        adj = ((torch.rand(B, N, N, generator=g) < 0.4) | torch.eye(N, dtype=torch.bool)[None]).to(dtype)
        edge_mask = adj[..., None].repeat(1, 1, 1, nh)

    h_in = h.clone()
    h_in._keras_mask = mask
    keras.backend.set_learning_phase(training)
    tf.random.seed(9000 + idx)
    model.transform_embeddings(h_in, e, edge_mask)          # reference layer loop (1 layer) + final norm
    keras.backend.set_learning_phase(0)
    L = model.tracked_layers.get_layers_dict()
    h_att = L['res_mha_00'].last_output
    e_att = L['res_edge_00'].last_output if 'res_edge_00' in L else e
    if cfg.add_n_norm:
        h_att = L['norm_mha_00'].last_output
        if 'norm_edge_00' in L:
            e_att = L['norm_edge_00'].last_output
    V_att, H_hat, A_tild = L['mha_00'].last_output
    h_ffn = L['res_fnn_node_00'].last_output
    e_ffn = L['res_fnn_edge_00'].last_output if 'res_fnn_edge_00' in L else None
    if cfg.add_n_norm:
        h_ffn = L['norm_fnn_node_00'].last_output
        if e_ffn is not None:
            e_ffn = L['norm_fnn_edge_00'].last_output
    draws = dict(tf.random.draws)

    # weights by reference layer name
    P = {}
    keymap = {'norm_mha_00': 'norm_mha', 'dense_qkv_00': 'dense_qkv', 'dense_mha_00': 'dense_mha',
              'norm_edge_00': 'norm_edge', 'attention_gates_00': 'attention_gates',
              'dense_edge_b_00': 'dense_edge_b', 'dense_edge_r_00': 'dense_edge_r',
              'norm_fnn_node_00': 'ffn_node/norm', 'fnn_lr1_node_00': 'ffn_node/lr1', 'fnn_lr2_node_00': 'ffn_node/lr2',
              'norm_fnn_edge_00': 'ffn_edge/norm', 'fnn_lr1_edge_00': 'ffn_edge/lr1', 'fnn_lr2_edge_00': 'ffn_edge/lr2'}
    for lname, short in keymap.items():
        if lname in L:
            for wn, w in L[lname]._weights.items():
                P[f'{short}/{wn}'] = w

    ocfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, gate_attention=cfg.gate_attention,
                         add_n_norm=cfg.add_n_norm,
                         clip_logits_value=None if cfg.clip_logits_value is None else tuple(cfg.clip_logits_value),
                         edge_activation=cfg.edge_activation, edge_channel_type=cfg.edge_channel_type,
                         scale_degree=cfg.scale_degree, scaler_type=cfg.scaler_type,
                         num_virtual_nodes=cfg.num_virtual_nodes, random_mask_prob=cfg.random_mask_prob,
                         attn_dropout=cfg.attn_dropout, activation=cfg.activation,
                         ffn_multiplier=cfg.ffn_multiplier)
    noise = {}
    if 'uniform' in draws:
        noise['random_mask'] = draws['uniform']
    o_h, o_e, aux = O.egt_block(h, e, mask, P, ocfg, edge_mask=edge_mask, training=bool(training),
                                noise=noise, return_aux=True)
    for a, b, nm in ((h_att, o_h, "h'"), (e_att, o_e, "e'"), (H_hat, aux['H_hat'], 'H_hat')):
        err = (a - b).abs().max().item()
        assert err < 1e-11, f'block case {idx} {nm}: restatement differs from reference source by {err}'
    # FFN half ("next" row)
    o_hf = O.ffn_channel(o_h, P, 'ffn_node', ocfg)
    assert (o_hf - h_ffn).abs().max().item() < 1e-11
    if e_ffn is not None:
        o_ef = O.ffn_channel(o_e, P, 'ffn_edge', ocfg)
        assert (o_ef - e_ffn).abs().max().item() < 1e-11

    rec = dict(h=_np(h), e=_np(e), mask=_np(mask), h_out=_np(h_att), e_out=_np(e_att),
               H_hat=_np(H_hat), V_att=_np(V_att), h_ffn=_np(h_ffn), training=np.int64(training))
    if e_ffn is not None:
        rec['e_ffn'] = _np(e_ffn)
    if edge_mask is not None:
        rec['edge_mask'] = _np(edge_mask)
    if 'random_mask' in noise:
        rec['uniform_noise'] = _np(noise['random_mask'])
    for k, v in P.items():
        rec['param:' + k] = _np(v)
    for k, v in vars(ocfg).items():
        if k == 'clip_logits_value':
            rec['cfg_clip'] = np.array([np.nan, np.nan] if v is None else v, dtype=np.float64)
        elif v is None:
            rec['cfg_' + k] = np.array('None')
        else:
            rec['cfg_' + k] = np.array(v)
    return rec


def main():
    torch.set_default_dtype(torch.float64)
    os.makedirs(OUT, exist_ok=True)
    lc = layer_cases()
    blob = {}
    for i, c in enumerate(lc):
        rec = run_layer_case(i, c)
        for k, v in rec.items():
            blob[f'case{i:02d}/{k}'] = v
    blob['n_cases'] = np.int64(len(lc))
    np.savez_compressed(os.path.join(OUT, 'egt_layer_reference.npz'), **blob)
    print(f'egt_layer_reference.npz: {len(lc)} cases')

    bc = block_cases()
    blob = {}
    for i, c in enumerate(bc):
        rec = run_block_case(i, c)
        for k, v in rec.items():
            blob[f'case{i:02d}/{k}'] = v
    blob['n_cases'] = np.int64(len(bc))
    np.savez_compressed(os.path.join(OUT, 'egt_block_reference.npz'), **blob)
    print(f'egt_block_reference.npz: {len(bc)} cases')


if __name__ == '__main__':
    main()
