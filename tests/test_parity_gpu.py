"""GPU parity: every test calls the CUDA path through the C ABI (egt_b200.ops -> libegt_b200.so)
and checks it against the CPU oracle / the golden vectors produced from the reference's source.

Tolerances (BASELINE.json north_star): fp32 rtol 1e-3, bf16 rtol 1e-2; mask / index behaviour
(which attention entries are exactly zero, scaler-1 rows, padded-key exclusion) bit-exact."""
import numpy as np
import pytest
import torch

from oracle import egt_oracle as O
from tests import philox
from tests.golden_util import layer_case, block_case

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
TOL = {torch.float32: dict(rtol=1e-3, atol=2e-4), torch.bfloat16: dict(rtol=1e-2, atol=1e-2)}


def _close(got, ref, dtype, what='', scale_atol=True):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    tol = TOL[dtype]
    atol = tol['atol'] * (max(1.0, float(ref.abs().max())) if scale_atol else 1.0)
    err = (got - ref).abs()
    bound = atol + tol['rtol'] * ref.abs()
    bad = err > bound
    assert not bad.any(), (f'{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; '
                           f'max err {float(err.max()):.3e} at ref {float(ref.flatten()[err.argmax()]):.3e}')


def _round(x, dtype):
    """The values the GPU actually sees (bf16 rounding of the inputs is shared with the oracle)."""
    return x.to(dtype).double()


# ------------------------------------------------------------------------------------------
# EGT layer  (egt_attn_fwd / egt_attn_bwd)  vs golden vectors and oracle
# ------------------------------------------------------------------------------------------
def _rows_with_a_live_key(c, flags, noise, B, N, h):
    """[B,N,h] bool: the query row has at least one key that no mask removes.  Rows without one
    (possible for padded queries under an attention mask, or under heavy random masking) are the
    appendix B-3 corner: only fp32 arithmetic reproduces TF there (x + -1e9 == -1e9)."""
    live = c['mask'][:, None, :, None].expand(B, N, N, h).clone()
    if flags['attn_mask']:
        live &= c['M'] > 0.5
    if noise is not None:
        live &= ~(noise.float() < flags['random_mask_prob'])
    return live.any(dim=2)


def _run_layer(c, dtype, seed=77, offset=5, want_grads=False):
    """GPU layer outputs + oracle outputs.  The oracle runs in fp64 on the inputs rounded to the GPU
    dtype; rows whose keys are ALL masked are taken from an fp32 oracle run instead."""
    import egt_b200
    flags = dict(c['flags'])
    layer = egt_b200.EGT(return_attn=True, seed=seed, **flags)
    layer.rng.offset = offset - 1
    nd = 1 + int(flags['edge_input']) + int(flags['gate_input'])      # differentiable inputs (not M)
    ins_gpu = [t.to(dtype).to(DEV).requires_grad_(want_grads and i < nd) for i, t in enumerate(c['inputs'])]
    mask = c['mask'].to(DEV)
    V, H, A = layer(ins_gpu, mask=mask, training=c['training'])
    B, N, _ = c['QKV'].shape
    h = flags['num_heads']
    kw, noise = {}, None
    if c['training'] and flags['random_mask_prob'] > 0:
        noise = torch.from_numpy(philox.noise_tensor(seed, offset, 0, B, N, h))
        kw['uniform_noise'] = noise
    if c['training'] and flags['attn_dropout'] > 0:
        kw['dropout_noise'] = torch.from_numpy(philox.noise_tensor(seed, offset, 1, B, N, h))
    live = _rows_with_a_live_key(c, flags, noise, B, N, h)            # [B,N,h]

    def oracle(odt):
        ins = [_round(t, dtype).to(odt).requires_grad_(want_grads) for t in c['inputs']]
        out = O.egt_layer(ins, mask=c['mask'], training=c['training'], **flags,
                          **{k: v.to(odt) for k, v in kw.items()})
        return ins, out

    ins64, (V64, H64, A64) = oracle(torch.float64)
    if bool(live.all()):
        return (V, H, A), (V64, H64, A64), ins_gpu, ins64, live
    ins32, (V32, H32, A32) = oracle(torch.float32)
    dk = V64.shape[-1] // h
    live_v = live[:, :, None, :].expand(B, N, dk, h).reshape(B, N, dk * h)
    Vr = torch.where(live_v, V64, V32.double())
    Ar = torch.where(live[:, :, None, :], A64, A32.double())
    return (V, H, A), (Vr, H64, Ar), ins_gpu, ins64, live


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('idx', range(19))
def test_layer_forward_golden(idx, dtype, golden_dir):
    c = layer_case(golden_dir, idx)
    (V, H, A), (Vr, Hr, Ar), _, _, live = _run_layer(c, dtype)
    _close(V, Vr, dtype, f'layer case {idx} V_att')
    _close(H, Hr, dtype, f'layer case {idx} H_hat')
    _close(A, Ar, dtype, f'layer case {idx} A_tild')
    # bit-exact mask behaviour: exactly the same entries are exactly zero
    assert torch.equal(A.cpu() == 0, Ar == 0), f'layer case {idx}: zero pattern of A_tild differs'
    if not c['training']:
        # deterministic cases must also match the committed golden outputs (made from the reference source)
        if dtype == torch.float32:
            B, N, h = live.shape
            lv = live[:, :, None, :].expand(B, N, V.shape[-1] // h, h).reshape(B, N, -1)
            _close(torch.where(lv, V.cpu().double(), c['V_att']), c['V_att'], dtype, f'layer case {idx} V_att vs golden')
            _close(H, c['H_hat'], dtype, f'layer case {idx} H_hat vs golden')
            assert torch.equal(A.cpu() == 0, c['A_tild'] == 0)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('idx', [0, 1, 2, 3, 5, 7, 8, 9, 10, 11, 12, 13, 14, 16])
def test_layer_backward(idx, dtype, golden_dir):
    c = layer_case(golden_dir, idx)
    (V, H, A), (Vr, Hr, Ar), ins_gpu, ins_o, live = _run_layer(c, dtype, want_grads=True)
    if not bool(live.all()):
        pytest.skip('all-keys-masked rows: forward corner is covered by test_layer_forward_golden / test_all_masked_rows')
    g = torch.Generator().manual_seed(idx)
    dV = torch.randn(V.shape, generator=g)
    dH = torch.randn(H.shape, generator=g)
    dVg, dHg = dV.to(dtype).to(DEV), dH.to(dtype).to(DEV)
    nd = 1 + int(c['flags']['edge_input']) + int(c['flags']['gate_input'])
    grads = torch.autograd.grad([V, H], ins_gpu[:nd], [dVg, dHg])
    ref = torch.autograd.grad([Vr, Hr], ins_o[:nd], [dVg.cpu().to(Vr.dtype), dHg.cpu().to(Hr.dtype)])
    for nm, a, b in zip(('dQKV', 'dE', 'dG'), grads, ref):
        _close(a, b, dtype, f'layer case {idx} {nm}')


def test_layer_errors():
    import egt_b200
    with pytest.raises(ValueError):
        egt_b200.EGT(scale_degree=True, gate_input=False)           # egt_layers.py:20-21
    with pytest.raises(ValueError):
        egt_b200.EGT(scaler_type='sqrt')                            # egt_layers.py:23-24
    layer = egt_b200.EGT(num_heads=8)
    q = torch.zeros(1, 4, 3 * 8 * 2 + 1, device=DEV)
    with pytest.raises(AssertionError):                             # egt_layers.py:70
        layer([q, torch.zeros(1, 4, 4, 8, device=DEV), torch.zeros(1, 4, 4, 8, device=DEV)])
    with pytest.raises(RuntimeError):                               # no CPU fallback
        layer([torch.zeros(1, 4, 48), torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 8)])


# ------------------------------------------------------------------------------------------
# attention block  (egt_block_fwd / egt_block_bwd)
# ------------------------------------------------------------------------------------------
def _spec_kwargs(cfg: O.BlockConfig):
    return dict(model_width=cfg.model_width, edge_width=cfg.edge_width, num_heads=cfg.num_heads,
                gate_attention=cfg.gate_attention, add_n_norm=cfg.add_n_norm,
                clip_logits_value=cfg.clip_logits_value, edge_activation=cfg.edge_activation,
                edge_channel_type=cfg.edge_channel_type, scale_degree=cfg.scale_degree,
                scaler_type=cfg.scaler_type, num_virtual_nodes=cfg.num_virtual_nodes,
                random_mask_prob=cfg.random_mask_prob, attn_dropout=cfg.attn_dropout)


def _oracle_params(block):
    return {k.replace('_00/', '/'): v.double() for k, v in block.keras_weights().items()}


def _run_block(cfg, params, h, e, mask, edge_mask, training, dtype, seed=31, offset=9, grads=False, force_path=None):
    import egt_b200
    blk = egt_b200.EGTBlock(seed=seed, **_spec_kwargs(cfg))
    blk.load_keras_weights({k: v for k, v in params.items() if not k.startswith('ffn')})
    blk = blk.to(DEV)
    blk.rng.offset = offset - 1
    hg = h.to(dtype).to(DEV).requires_grad_(grads)
    eg = e.to(dtype).to(DEV).requires_grad_(grads)
    emg = None if edge_mask is None else edge_mask.to(DEV)
    h2, e2 = blk(hg, eg, mask.to(DEV), edge_mask=emg, training=training)
    B, N, _ = h.shape
    noise = None
    if training and cfg.random_mask_prob > 0:
        noise = {'random_mask': torch.from_numpy(philox.noise_tensor(seed, offset, 0, B, N, cfg.num_heads)).double()}
    hr = _round(h, dtype).requires_grad_(grads)
    er = _round(e, dtype).requires_grad_(grads)
    pr = {k: v.double().clone().requires_grad_(grads) for k, v in params.items() if not k.startswith('ffn')}
    h2r, e2r = O.egt_block(hr, er, mask, pr, cfg, edge_mask=None if edge_mask is None else edge_mask.double(),
                           training=training, noise=noise)
    return blk, (hg, eg, h2, e2), (hr, er, pr, h2r, e2r)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('idx', range(16))
def test_block_forward_golden(idx, dtype, golden_dir):
    c = block_case(golden_dir, idx)
    if c['cfg'].add_n_norm:
        import egt_b200
        with pytest.raises(NotImplementedError):
            egt_b200.EGTBlock(**_spec_kwargs(c['cfg']))
        return
    blk, (hg, eg, h2, e2), (hr, er, pr, h2r, e2r) = _run_block(
        c['cfg'], c['params'], c['h'], c['e'], c['mask'], c.get('edge_mask'), c['training'], dtype)
    _close(h2, h2r, dtype, f"block case {idx} h'")
    _close(e2, e2r, dtype, f"block case {idx} e'")
    if dtype == torch.float32 and not c['training']:
        _close(h2, c['h_out'], dtype, f"block case {idx} h' vs golden")
        _close(e2, c['e_out'], dtype, f"block case {idx} e' vs golden")


BWD_CASES = [
    # (B, N, d, d_e, h, overrides)
    (2, 9, 16, 8, 4, {}),
    (2, 9, 16, 8, 4, dict(edge_channel_type='bias')),
    (2, 9, 16, 8, 4, dict(edge_channel_type='none')),
    (2, 9, 16, 8, 4, dict(edge_channel_type='constrained', scale_degree=True)),
    (2, 9, 16, 8, 4, dict(gate_attention=False)),
    (2, 9, 16, 8, 4, dict(scale_degree=True, scaler_type='linear', num_virtual_nodes=1)),
    (2, 9, 16, 8, 4, dict(edge_activation='lrelu2')),
    (2, 9, 16, 8, 4, dict(edge_activation='elu', clip_logits_value=None)),
    (2, 9, 16, 8, 4, dict(random_mask_prob=0.2, training=True, scale_degree=True)),
    (3, 37, 64, 64, 8, dict(scale_degree=True)),                 # ZINC-500K widths, N=37
    (2, 13, 48, 48, 8, {}),                                      # ZINC-100K widths, dk=6
    (2, 21, 96, 8, 8, dict(scale_degree=True)),                  # BASELINE CLUSTER d=96, dk=12
    (2, 19, 128, 32, 16, dict(scale_degree=True)),               # C5 widths
    (2, 75, 64, 8, 8, dict(random_mask_prob=0.1, training=True)),  # MNIST shape, shipped training flags
    (1, 130, 64, 8, 8, {}),                                      # N just above one 128-row tile
]


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('case', range(len(BWD_CASES)))
def test_block_forward_backward_vs_oracle(case, dtype):
    B, N, d, de, nh, ov = BWD_CASES[case]
    ov = dict(ov)
    training = ov.pop('training', False)
    cfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, **ov)
    params = O.init_block_params(cfg, seed=1234 + case, dtype=torch.float64)
    h, e, mask = O.synthetic_batch(B, N, d, de, seed=20240 + case, ragged=True, dtype=torch.float64)
    edge_mask = None
    if cfg.edge_channel_type == 'constrained':
        g = torch.Generator().manual_seed(case)
        adj = ((torch.rand(B, N, N, generator=g) < 0.4) | torch.eye(N, dtype=torch.bool)[None]).double()
        edge_mask = adj[..., None].repeat(1, 1, 1, nh)
    blk, (hg, eg, h2, e2), (hr, er, pr, h2r, e2r) = _run_block(cfg, params, h, e, mask, edge_mask, training, dtype,
                                                               grads=True)
    _close(h2, h2r, dtype, "h'")
    _close(e2, e2r, dtype, "e'")
    g = torch.Generator().manual_seed(99 + case)
    dh = torch.randn(h2.shape, generator=g).to(dtype)
    de_ = torch.randn(e2.shape, generator=g).to(dtype)
    residual = cfg.edge_channel_type in ('residual', 'constrained')
    if residual:
        gin = torch.autograd.grad([h2, e2], [hg, eg, blk.flat], [dh.to(DEV), de_.to(DEV)], allow_unused=True)
        rin = torch.autograd.grad([h2r, e2r], [hr, er] + list(pr.values()), [dh.double(), de_.double()],
                                  allow_unused=True)
    else:
        gin = torch.autograd.grad([h2], [hg, eg, blk.flat], [dh.to(DEV)], allow_unused=True)
        rin = torch.autograd.grad([h2r], [hr, er] + list(pr.values()), [dh.double()], allow_unused=True)
    _close(gin[0], rin[0], dtype, 'dh')
    if cfg.edge_channel_type != 'none':
        _close(gin[1], rin[1], dtype, 'de')
    blk.flat.grad = gin[2]
    for (name, _), gr in zip(pr.items(), rin[2:]):
        field = name.replace('/', '_')
        got = blk.grad_view(field)
        # weight gradients are sums over B*N^2 terms: compare relative to their own scale
        tol = 1e-3 if dtype == torch.float32 else 1e-2   # north_star: 1e-3 fp32 / 1e-2 bf16, relative to the tensor's own maximum
        # a bias gradient can be analytically ~0 (e.g. dense_edge_b/bias in 'bias' mode: softmax is
        # shift-invariant), so its error is measured against the scale of its kernel's gradient (a sum of
        # B*N^2 bf16-rounded terms carries the same absolute noise as the kernel's gradient does)
        kname = name.rsplit('/', 1)[0] + '/kernel'
        floor = 1e-1 * float(dict(zip(pr, rin[2:]))[kname].abs().max()) if name.endswith('bias') and kname in pr else 0.
        denom = max(float(gr.abs().max()), floor, 1e-6)
        err = float((got.double().cpu() - gr).abs().max()) / denom
        # when the floor is what the error is measured against (a gradient that is analytically ~0), the noise of
        # the ~B*N^2 bf16-rounded terms is not relative to the tensor itself: twice the tolerance
        tol_t = 2 * tol if floor > float(gr.abs().max()) else tol
        assert err < tol_t, f'grad {name}: rel-to-max err {err:.3e}'


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('case', [0, 3, 9, 10, 11, 12])
def test_any_shape_staged_kernels_vs_oracle(case, dtype, monkeypatch):
    """The shape-specialised staged kernels (edge_fast.cu, attn_fast.cu) take the common (d_e, h, dk); the
    any-shape kernels behind them (edge_kernels.cu, attn_staged.cu) must stay correct on those shapes too."""
    monkeypatch.setenv('EGT_STAGED_GENERIC', '1')
    test_block_forward_backward_vs_oracle(case, dtype)


def test_block_mask_is_bit_exact_and_padding_is_inert():
    """Padded KEYS never influence valid rows: perturbing padded nodes / padded edge columns leaves
    h' and e' at valid positions bit-identical (egt_layers.py:91-94)."""
    import egt_b200
    torch.manual_seed(0)
    B, N, d, de, nh = 4, 50, 64, 8, 8
    blk = egt_b200.EGTBlock(model_width=d, edge_width=de, num_heads=nh, scale_degree=True).to(DEV)
    nn_ = torch.tensor([50, 33, 17, 1])
    mask = (torch.arange(N)[None] < nn_[:, None]).to(DEV)
    h = torch.randn(B, N, d, device=DEV).bfloat16()
    e = torch.randn(B, N, N, de, device=DEV).bfloat16()
    h2, e2 = blk(h, e, mask)
    hp, ep = h.clone(), e.clone()
    for b in range(B):
        n = int(nn_[b])
        hp[b, n:] = torch.randn(N - n, d, device=DEV).bfloat16() * 3
        ep[b, :, n:] = torch.randn(N, N - n, de, device=DEV).bfloat16() * 3
    h2p, e2p = blk(hp, ep, mask)
    for b in range(B):
        n = int(nn_[b])
        assert torch.equal(h2[b, :n], h2p[b, :n])
        assert torch.equal(e2[b, :n, :n], e2p[b, :n, :n])
    assert torch.isfinite(h2.float()).all() and torch.isfinite(e2.float()).all()   # padded rows: finite garbage


@pytest.mark.parametrize('N,d,de,nh,B', [(128, 64, 8, 8, 16), (190, 96, 8, 8, 4), (188, 64, 8, 8, 4),
                                         (75, 64, 8, 8, 16), (37, 64, 64, 8, 16), (512, 128, 32, 16, 2),
                                         (512, 64, 8, 8, 3), (64, 64, 8, 8, 16), (256, 64, 8, 8, 3)])
def test_block_full_size_vs_oracle_sample(N, d, de, nh, B):
    """BASELINE.json shapes at full N and the N = 64 / 256 / 512 points of the sweep on the fused path: oracle on two
    graphs of the batch + permutation equivariance."""
    import egt_b200
    cfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, scale_degree=True)
    params = O.init_block_params(cfg, dtype=torch.float64)
    h, e, mask = O.synthetic_batch(B, N, d, de, ragged=True, dtype=torch.float32)
    blk = egt_b200.EGTBlock(**_spec_kwargs(cfg))
    blk.load_keras_weights(params)
    blk = blk.to(DEV)
    hg, eg = h.bfloat16().to(DEV).requires_grad_(True), e.bfloat16().to(DEV).requires_grad_(True)
    h2, e2 = blk(hg, eg, mask.to(DEV))
    g = torch.Generator().manual_seed(1)
    dh, de_ = torch.randn(h2.shape, generator=g).bfloat16(), torch.randn(e2.shape, generator=g).bfloat16()
    gh, ge = torch.autograd.grad([h2, e2], [hg, eg], [dh.to(DEV), de_.to(DEV)])
    sel = [0, B - 1]
    hr = _round(h[sel], torch.bfloat16).requires_grad_(True)
    er = _round(e[sel], torch.bfloat16).requires_grad_(True)
    h2r, e2r = O.egt_block(hr, er, mask[sel], params, cfg)
    ghr, ger = torch.autograd.grad([h2r, e2r], [hr, er], [dh[sel].double(), de_[sel].double()])
    _close(h2[sel], h2r, torch.bfloat16, "h'")
    _close(e2[sel], e2r, torch.bfloat16, "e'")
    _close(gh[sel], ghr, torch.bfloat16, 'dh')
    _close(ge[sel], ger, torch.bfloat16, 'de')
    # weight gradients at full size: the same two graphs as their own batch, every tensor against the oracle
    pr = {k: v.double().clone().requires_grad_(True) for k, v in params.items() if not k.startswith('ffn')}
    hr2 = _round(h[sel], torch.bfloat16)
    er2 = _round(e[sel], torch.bfloat16)
    h2w, e2w = O.egt_block(hr2, er2, mask[sel], pr, cfg)
    wref = torch.autograd.grad([h2w, e2w], list(pr.values()), [dh[sel].double(), de_[sel].double()])
    hs, es = h[sel].bfloat16().to(DEV).requires_grad_(True), e[sel].bfloat16().to(DEV).requires_grad_(True)
    blk.flat.grad = None
    h2s, e2s = blk(hs, es, mask[sel].to(DEV))
    torch.autograd.backward([h2s, e2s], [dh[sel].to(DEV), de_[sel].to(DEV)])
    for (name, _), gr in zip(pr.items(), wref):
        got = blk.grad_view(name.replace('/', '_'))
        err = float((got.double().cpu() - gr).abs().max()) / max(float(gr.abs().max()), 1e-6)
        assert err < 1e-2, f'full-size weight gradient {name}: rel-to-max err {err:.3e}'
    # graphs are independent: permuting the batch permutes the outputs bit-exactly
    perm = torch.randperm(B, generator=g)
    h2p, e2p = blk(hg.detach()[perm], eg.detach()[perm], mask[perm].to(DEV))
    assert torch.equal(h2p, h2.detach()[perm]) and torch.equal(e2p, e2.detach()[perm])


def test_gradient_linearity():
    """backward is linear in the upstream gradients (size-independent property)."""
    import egt_b200
    torch.manual_seed(1)
    B, N, d, de, nh = 8, 64, 64, 8, 8
    blk = egt_b200.EGTBlock(model_width=d, edge_width=de, num_heads=nh).to(DEV)
    h = torch.randn(B, N, d, device=DEV, requires_grad=True)
    e = torch.randn(B, N, N, de, device=DEV, requires_grad=True)
    h2, e2 = blk(h, e, None)
    a, b = torch.randn_like(h2), torch.randn_like(e2)
    g1 = torch.autograd.grad([h2, e2], [h, e, blk.flat], [a, b], retain_graph=True)
    g2 = torch.autograd.grad([h2, e2], [h, e, blk.flat], [2 * a, 2 * b], retain_graph=True)
    for x, y in zip(g1, g2):
        # fp32 weight gradients are accumulated with atomics (order varies run to run): fp32 tolerance of the path
        torch.testing.assert_close(2 * x, y, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize('gated', [False, True])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_all_masked_rows(gated, dtype):
    """Appendix B-3: a row whose keys are all masked softmaxes uniformly over the least-masked keys
    (ungated output = mean of those V rows; gated output = 0), forward and backward."""
    import egt_b200
    B, N, h, dk = 2, 6, 4, 4
    g = torch.Generator().manual_seed(5)
    QKV = torch.randn(B, N, 3 * h * dk, generator=g)
    E = torch.randn(B, N, N, h, generator=g)
    G = torch.randn(B, N, N, h, generator=g)
    M = torch.ones(B, N, N, h)
    M[0, 2] = 0          # row 2 of graph 0 may attend to nothing
    M[1, :, 1:] = 0      # graph 1: only key 0 allowed ...
    mask = torch.ones(B, N, dtype=torch.bool)
    mask[1, 0] = False   # ... and key 0 is padding -> every row of graph 1 is fully masked
    layer = egt_b200.EGT(num_heads=h, gate_input=gated, attn_mask=True, return_attn=True)
    ins = [QKV, E] + ([G] if gated else []) + [M]
    nd = 3 if gated else 2
    ins_g = [t.to(dtype).to(DEV).requires_grad_(i < nd) for i, t in enumerate(ins)]
    V, H, A = layer(ins_g, mask=mask.to(DEV))
    ins_o = [t.to(dtype).float().requires_grad_(i < nd) for i, t in enumerate(ins)]      # fp32 = TF semantics
    Vr, Hr, Ar = O.egt_layer(ins_o, mask=mask, num_heads=h, gate_input=gated, attn_mask=True)
    _close(V, Vr, dtype, 'V_att')
    _close(A, Ar, dtype, 'A_tild')
    assert torch.equal(A.cpu() == 0, Ar == 0)
    dV = torch.randn(V.shape, generator=g).to(dtype)
    dH = torch.randn(H.shape, generator=g).to(dtype)
    gg = torch.autograd.grad([V, H], ins_g[:nd], [dV.to(DEV), dH.to(DEV)])
    gr = torch.autograd.grad([Vr, Hr], ins_o[:nd], [dV.float(), dH.float()])
    for a, b, nm in zip(gg, gr, ('dQKV', 'dE', 'dG')):
        _close(a, b, dtype, nm)


@pytest.mark.parametrize('training,rmp', [(False, 0.), (True, 0.1)])
@pytest.mark.parametrize('N,B', [(128, 6), (75, 5), (37, 7), (190, 3), (9, 4), (1, 3), (257, 2)])
def test_fused_path_matches_staged_and_oracle(N, B, training, rmp):
    """Shapes of the reference's MNIST/CLUSTER/PATTERN configs (d=64, d_e=8, h=8) dispatch to the fused
    tcgen05 kernels; they must agree with the staged kernels and the oracle (incl. the same random key mask)."""
    import egt_b200
    from egt_b200 import _lib as L
    lib = L.load()
    d, de, nh = 64, 8, 8
    cfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, scale_degree=True, random_mask_prob=rmp)
    params = O.init_block_params(cfg, seed=5, dtype=torch.float64)
    h, e, mask = O.synthetic_batch(B, N, d, de, seed=11, ragged=True, dtype=torch.float64)
    outs = {}
    for force in (0, 1):
        lib.egt_debug_force_staged(force)
        try:
            blk, (hg, eg, h2, e2), (hr, er, pr, h2r, e2r) = _run_block(cfg, params, h, e, mask, None, training,
                                                                       torch.bfloat16, grads=True)
            assert lib.egt_last_path() == (0 if force else 1)
            g = torch.Generator().manual_seed(3)
            dh = torch.randn(h2.shape, generator=g).bfloat16()
            de_ = torch.randn(e2.shape, generator=g).bfloat16()
            gin = torch.autograd.grad([h2, e2], [hg, eg, blk.flat], [dh.to(DEV), de_.to(DEV)])
            rin = torch.autograd.grad([h2r, e2r], [hr, er], [dh.double(), de_.double()])
        finally:
            lib.egt_debug_force_staged(0)
        _close(h2, h2r, torch.bfloat16, f"h' force_staged={force}")
        _close(e2, e2r, torch.bfloat16, f"e' force_staged={force}")
        _close(gin[0], rin[0], torch.bfloat16, f'dh force_staged={force}')
        _close(gin[1], rin[1], torch.bfloat16, f'de force_staged={force}')
        outs[force] = (h2, e2, gin)
    # weight gradients of the two paths agree to bf16 accumulation noise
    wa, wb = outs[0][2][2], outs[1][2][2]
    assert float((wa - wb).abs().max()) <= 2e-2 * float(wb.abs().max())   # two bf16 paths: twice the 1e-2 of either against the oracle


def test_block_step_replays_as_cuda_graph():
    """bench.py replays the step as a CUDA graph: a captured forward+backward must give the same result as
    the eager launches when new data is copied into the static input buffers."""
    import egt_b200
    torch.manual_seed(3)
    B, N, d, de, nh = 6, 96, 64, 8, 8
    blk = egt_b200.EGTBlock(model_width=d, edge_width=de, num_heads=nh, scale_degree=True).to(DEV)
    with torch.no_grad():
        blk.flat.add_(0.05 * torch.randn_like(blk.flat))
    mask = (torch.arange(N)[None] < torch.tensor([96, 80, 64, 50, 96, 7])[:, None]).to(DEV)
    hs = torch.zeros(B, N, d, device=DEV, dtype=torch.bfloat16)
    es = torch.zeros(B, N, N, de, device=DEV, dtype=torch.bfloat16)
    dh = torch.randn(B, N, d, device=DEV).bfloat16()
    dE = torch.randn(B, N, N, de, device=DEV).bfloat16()

    def step(h, e):
        h = h.detach().requires_grad_(True)
        e = e.detach().requires_grad_(True)
        blk.flat.grad = None
        h2, e2 = blk(h, e, mask)
        torch.autograd.backward([h2, e2], [dh, dE])
        return h2.detach(), e2.detach(), h.grad, e.grad, blk.flat.grad

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(hs, es)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = step(hs, es)
    for trial in range(2):
        h = torch.randn(B, N, d, device=DEV).bfloat16()
        e = torch.randn(B, N, N, de, device=DEV).bfloat16()
        hs.copy_(h)
        es.copy_(e)
        g.replay()
        torch.cuda.synchronize()
        got = [o.clone() for o in outs]
        ref = step(h, e)
        for a, b, nm in zip(got[:4], ref[:4], ('h2', 'e2', 'dh', 'de')):
            assert torch.equal(a, b), f'{nm} differs between graph replay and eager launches (trial {trial})'
        # weight gradients are accumulated with atomics across CTAs: equal up to summation order
        torch.testing.assert_close(got[4], ref[4], rtol=1e-3, atol=1e-3 * float(ref[4].abs().max()))


def test_fused_forward_then_backward_without_edge_gradient():
    """When nothing downstream uses e', autograd materialises a zero gradient for it and the fused backward
    runs with de' = 0 (the C ABI additionally accepts de_out = NULL and then runs the staged kernels on what
    the fused forward saved: pre-scaled Q, log row sums with a zero reference)."""
    import egt_b200
    from egt_b200 import _lib as L
    lib = L.load()
    B, N, d, de, nh = 3, 70, 64, 8, 8
    cfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, scale_degree=True)
    params = O.init_block_params(cfg, seed=9, dtype=torch.float64)
    h, e, mask = O.synthetic_batch(B, N, d, de, seed=4, ragged=True, dtype=torch.float64)
    blk, (hg, eg, h2, e2), (hr, er, pr, h2r, e2r) = _run_block(cfg, params, h, e, mask, None, False, torch.bfloat16,
                                                               grads=True)
    assert lib.egt_last_path() == 1
    g = torch.Generator().manual_seed(3)
    dh = torch.randn(h2.shape, generator=g).bfloat16()
    gin = torch.autograd.grad([h2], [hg, eg], [dh.to(DEV)])
    rin = torch.autograd.grad([h2r], [hr, er], [dh.double()])
    _close(gin[0], rin[0], torch.bfloat16, 'dh')
    _close(gin[1], rin[1], torch.bfloat16, 'de')
