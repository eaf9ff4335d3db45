"""CPU: the oracle (oracle/egt_oracle.py) is pinned against tests/golden/*.npz, which were produced
by executing the reference's OWN source files (egt_layers.py, graph_xformer_model_base.py) on the
TF shim (oracle/make_golden.py).  Also pins the hand-derived backward against autograd."""
import os

import numpy as np
import pytest
import torch

from oracle import egt_oracle as O
from tests.golden_util import layer_case, block_case, n_cases


@pytest.mark.parametrize('idx', range(19))
def test_layer_golden(idx, golden_dir):
    c = layer_case(golden_dir, idx)
    V, H, A = O.egt_layer(c['inputs'], mask=c['mask'], training=c['training'],
                          uniform_noise=c.get('uniform_noise'), dropout_noise=c.get('dropout_noise'),
                          **c['flags'])
    for got, key in ((V, 'V_att'), (H, 'H_hat'), (A, 'A_tild')):
        np.testing.assert_allclose(got.numpy(), c[key].numpy(), rtol=0, atol=1e-12, err_msg=f'case {idx} {key}')


def test_golden_counts(golden_dir):
    assert n_cases(golden_dir, 'layer') == 19
    assert n_cases(golden_dir, 'block') == 16


@pytest.mark.parametrize('idx', range(16))
def test_block_golden(idx, golden_dir):
    c = block_case(golden_dir, idx)
    noise = {'random_mask': c['uniform_noise']} if 'uniform_noise' in c else None
    h2, e2, aux = O.egt_block(c['h'], c['e'], c['mask'], c['params'], c['cfg'], edge_mask=c.get('edge_mask'),
                              training=c['training'], noise=noise, return_aux=True)
    np.testing.assert_allclose(h2.numpy(), c['h_out'].numpy(), rtol=0, atol=1e-11)
    np.testing.assert_allclose(e2.numpy(), c['e_out'].numpy(), rtol=0, atol=1e-11)
    np.testing.assert_allclose(aux['H_hat'].numpy(), c['H_hat'].numpy(), rtol=0, atol=1e-11)
    # FFN half ("next" row 8f-1) is pinned too
    hf = O.ffn_channel(h2, c['params'], 'ffn_node', c['cfg'])
    np.testing.assert_allclose(hf.numpy(), c['h_ffn'].numpy(), rtol=0, atol=1e-11)


@pytest.mark.parametrize('ect,gated,scale,act', [
    ('residual', True, False, None), ('residual', True, True, None), ('residual', False, False, None),
    ('bias', True, True, 'lrelu2'), ('none', True, False, None), ('constrained', True, True, 'elu'),
    ('residual', True, True, 'tanh'),
])
def test_closed_form_backward_matches_autograd(ect, gated, scale, act):
    torch.manual_seed(0)
    dt = torch.float64
    cfg = O.BlockConfig(model_width=16, edge_width=8, num_heads=4, gate_attention=gated, scale_degree=scale and gated and ect != 'none',
                        edge_channel_type=ect, edge_activation=act, num_virtual_nodes=1 if scale else 0,
                        random_mask_prob=0.2)
    B, N = 2, 7
    p = O.init_block_params(cfg, dtype=dt)
    h, e, mask = O.synthetic_batch(B, N, 16, 8, ragged=True, dtype=dt)
    em = None
    if ect == 'constrained':
        adj = ((torch.rand(B, N, N) < 0.5) | torch.eye(N, dtype=torch.bool)[None]).to(dt)
        em = adj[..., None].repeat(1, 1, 1, 4)
    noise = {'random_mask': torch.rand(B, N, N, 4, dtype=dt)}
    hr, er = h.clone().requires_grad_(True), e.clone().requires_grad_(True)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    h2, e2 = O.egt_block(hr, er, mask, pr, cfg, edge_mask=em, training=True, noise=noise)
    dh_out, de_out = torch.randn_like(h2), torch.randn_like(e2)
    loss = (h2 * dh_out).sum() + (e2 * de_out).sum()
    names = list(pr)
    grads = torch.autograd.grad(loss, [hr, er] + [pr[k] for k in names], allow_unused=True)
    dh, de, gw = O.egt_block_backward(h, e, mask, p, cfg, dh_out, de_out, edge_mask=em, training=True, noise=noise)
    np.testing.assert_allclose(dh.numpy(), grads[0].numpy(), atol=1e-10)
    np.testing.assert_allclose(de.numpy(), grads[1].numpy(), atol=1e-10)
    for k, g in zip(names, grads[2:]):
        np.testing.assert_allclose(gw[k].numpy(), g.numpy(), atol=1e-10, err_msg=k)


def test_all_keys_masked_row_is_uniform_in_fp32():
    """in fp32 (what TF runs) x + (-1e9) == -1e9 for |x| < 32, so a row whose
    keys are all masked softmaxes to 1/N while its gates are exactly 0."""
    B, N, h, dk = 1, 5, 2, 2
    g = torch.Generator().manual_seed(3)
    QKV = torch.randn(B, N, 3 * h * dk, generator=g)
    E = torch.randn(B, N, N, h, generator=g)
    G = torch.randn(B, N, N, h, generator=g)
    noise = torch.zeros(B, N, N, h)          # u=0 < p: every key masked
    V, H, A = O.egt_layer([QKV, E, G], mask=torch.ones(B, N, dtype=torch.bool), training=True, num_heads=h,
                          random_mask_prob=0.5, uniform_noise=noise)
    assert torch.all(A == 0) and torch.all(V == 0)
    V2, _, A2 = O.egt_layer([QKV, E], mask=torch.ones(B, N, dtype=torch.bool), training=True, num_heads=h,
                            gate_input=False, random_mask_prob=0.5, uniform_noise=noise)
    np.testing.assert_allclose(A2.numpy(), np.full((B, N, N, h), 1.0 / N, dtype=np.float32), rtol=1e-6)


def test_mask_producers():
    feat = torch.tensor([[3, 1, -1, -1], [0, 2, 5, -1]])
    m = O.node_mask_from_features(feat)
    assert m.tolist() == [[True, True, False, False], [True, True, True, False]]
    m2 = O.prepend_virtual_nodes_mask(m, 2)
    assert m2.shape == (2, 6) and m2[:, :2].all()
    adj = torch.eye(4)[None].repeat(2, 1, 1)
    em = O.edge_mask_from_adjacency(adj, 3, n_vn=1)
    assert em.shape == (2, 5, 5, 3) and em[:, 0].all() and em[:, :, 0].all()
