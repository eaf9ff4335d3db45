"""Width-generic fused tcgen05 kernels (csrc/wide_fwd.cu, wide_bwd.cu) at the widths BASELINE.json names besides the
headline one: C5 / sweep (d=128, h=16, d_e=32), C1 ZINC (d=64, h=8, d_e=64), C3 CLUSTER (d=96, h=8, dk=12, d_e=8).
Every case runs through the C ABI, must dispatch to the fused path (egt_last_path() == 1) and is compared with the
oracle and with the staged CUDA-core kernels on the same inputs (incl. the same random key mask)."""
import pytest
import torch

from oracle import egt_oracle as O
from tests.test_parity_gpu import DEV, _close, _run_block

pytestmark = pytest.mark.gpu

# dense_edge_b scale of test_fused_large_edge_logits: max |E| ~ 300-500.  Every fused forward takes the exact row
# maximum from a pre-pass when the logit bound exceeds the fp32 exponent budget (DESIGN.md "softmax reference")
LARGE_WSCALE = {'C0': 100.0, 'C5': 60.0, 'C1': 60.0, 'C3': 60.0}
WIDTHS = {'C5': (128, 32, 16), 'C1': (64, 64, 8), 'C3': (96, 8, 8), 'C0': (64, 8, 8)}


def _case(width, N, B, training, rmp, seed=5, scale_degree=True, wscale=None):
    d, de, nh = WIDTHS[width]
    cfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, scale_degree=scale_degree, random_mask_prob=rmp)
    params = O.init_block_params(cfg, seed=seed, dtype=torch.float64)
    if wscale is not None:
        params['dense_edge_b/kernel'] = params['dense_edge_b/kernel'] * wscale
    h, e, mask = O.synthetic_batch(B, N, d, de, seed=11, ragged=True, dtype=torch.float64)
    return cfg, params, h, e, mask


def _fwd_bwd(cfg, params, h, e, mask, training, force):
    from egt_b200 import _lib as L
    lib = L.load()
    lib.egt_debug_force_staged(force)
    try:
        blk, (hg, eg, h2, e2), (hr, er, pr, h2r, e2r) = _run_block(cfg, params, h, e, mask, None, training,
                                                                   torch.bfloat16, grads=True)
        path_fwd = lib.egt_last_path()
        g = torch.Generator().manual_seed(3)
        dh = torch.randn(h2.shape, generator=g).bfloat16()
        de_ = torch.randn(e2.shape, generator=g).bfloat16()
        gin = torch.autograd.grad([h2, e2], [hg, eg, blk.flat], [dh.to(DEV), de_.to(DEV)])
        path_bwd = lib.egt_last_path()
        rin = torch.autograd.grad([h2r, e2r], [hr, er] + list(pr.values()), [dh.double(), de_.double()])
    finally:
        lib.egt_debug_force_staged(0)
    return blk, (h2, e2, gin), (h2r, e2r, rin, pr), (path_fwd, path_bwd)


@pytest.mark.parametrize('training,rmp', [(False, 0.), (True, 0.1)])
@pytest.mark.parametrize('N,B', [(37, 5), (128, 3), (9, 4), (1, 2), (200, 2), (130, 2)])
@pytest.mark.parametrize('width', ['C5', 'C1', 'C3'])
def test_wide_path_matches_staged_and_oracle(width, N, B, training, rmp):
    cfg, params, h, e, mask = _case(width, N, B, training, rmp)
    outs = {}
    for force in (0, 1):
        blk, (h2, e2, gin), (h2r, e2r, rin, pr), paths = _fwd_bwd(cfg, params, h, e, mask, training, force)
        assert paths[0] == (0 if force else 1), f'forward dispatched to path {paths[0]}'
        _close(h2, h2r, torch.bfloat16, f"h' force_staged={force}")
        _close(e2, e2r, torch.bfloat16, f"e' force_staged={force}")
        _close(gin[0], rin[0], torch.bfloat16, f'dh force_staged={force}')
        _close(gin[1], rin[1], torch.bfloat16, f'de force_staged={force}')
        outs[force] = gin[2]
        # weight gradients vs the oracle, relative to each tensor's own scale (north_star: 1e-2 for bf16)
        blk.flat.grad = gin[2]
        for (name, _), gr in zip(pr.items(), rin[2:]):
            got = blk.grad_view(name.replace('/', '_'))
            kname = name.rsplit('/', 1)[0] + '/kernel'
            floor = 1e-1 * float(dict(zip(pr, rin[2:]))[kname].abs().max()) if name.endswith('bias') and kname in pr else 0.
            denom = max(float(gr.abs().max()), floor, 1e-6)
            err = float((got.double().cpu() - gr).abs().max()) / denom
            assert err < 1e-2, f'grad {name}: rel-to-max err {err:.3e} (force_staged={force})'
    assert float((outs[0] - outs[1]).abs().max()) <= 2e-2 * float(outs[1].abs().max())


@pytest.mark.parametrize('width', ['C0', 'C5', 'C1', 'C3'])
def test_fused_large_edge_logits(width):
    """|E| of the order of 100: tf.nn.softmax (egt_layers.py:111) subtracts the row maximum; the fused kernels
    use the data-independent bound of the logits as the exponent reference instead and must agree (C0 = the
    headline widths on fused_fwd.cu / fused_bwd.cu)."""
    cfg, params, h, e, mask = _case(width, 50, 3, False, 0., wscale=LARGE_WSCALE[width])
    blk, (h2, e2, gin), (h2r, e2r, rin, pr), paths = _fwd_bwd(cfg, params, h, e, mask, False, 0)
    assert paths[0] == 1
    assert torch.isfinite(h2.float()).all() and torch.isfinite(e2.float()).all()
    _close(h2, h2r, torch.bfloat16, "h'")
    _close(e2, e2r, torch.bfloat16, "e'")
    _close(gin[0], rin[0], torch.bfloat16, 'dh')
    _close(gin[1], rin[1], torch.bfloat16, 'de')


@pytest.mark.parametrize('N,d,de,nh,B,ks', [(512, 128, 32, 16, 2, 4), (512, 128, 32, 16, 2, 8), (190, 96, 8, 8, 2, 2),
                                            (300, 64, 64, 8, 2, 4)])
def test_wide_backward_key_split(N, d, de, nh, B, ks, monkeypatch):
    """The wide backward splits the keys of a (graph, row tile) over several CTAs when that fills the last wave of the grid
    (csrc/wide_bwd.cu: wide_bwd_key_splits; dQ is then accumulated with atomics).  Forced here with EGT_WIDE_KSPLIT: every
    gradient equals the unsplit kernel's, and the oracle's on the sampled graphs."""
    from tests.test_parity_gpu import test_block_full_size_vs_oracle_sample, _spec_kwargs
    import egt_b200
    from oracle import egt_oracle as O
    cfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, scale_degree=True)
    params = O.init_block_params(cfg, dtype=torch.float64)
    h, e, mask = O.synthetic_batch(B, N, d, de, ragged=True, dtype=torch.float32)
    blk = egt_b200.EGTBlock(**_spec_kwargs(cfg))
    blk.load_keras_weights(params)
    blk = blk.to(DEV)
    g = torch.Generator().manual_seed(3)
    dh = torch.randn(B, N, d, generator=g).bfloat16().to(DEV)
    de_ = torch.randn(B, N, N, de, generator=g).bfloat16().to(DEV)

    def run():
        hg, eg = h.bfloat16().to(DEV).requires_grad_(True), e.bfloat16().to(DEV).requires_grad_(True)
        blk.flat.grad = None
        h2, e2 = blk(hg, eg, mask.to(DEV))
        torch.autograd.backward([h2, e2], [dh, de_])
        return hg.grad.clone(), eg.grad.clone(), blk.flat.grad.clone()

    monkeypatch.setenv('EGT_WIDE_KSPLIT', '1')
    ref = run()
    monkeypatch.setenv('EGT_WIDE_KSPLIT', str(ks))
    got = run()
    assert torch.equal(got[1], ref[1])                       # de: per (row, key), untouched by the split
    for a_, b_, what in zip(got, ref, ('dh', 'de', 'weight gradients')):
        scale = float(b_.float().abs().max())
        err = float((a_.float() - b_.float()).abs().max())
        # summation order differs (atomics): float32 weight gradients to 2e-3, bf16 dh to one bf16 step of its largest values
        tol = 8e-3 if a_.dtype == torch.bfloat16 else 2e-3
        assert err <= tol * scale, f'{what}: {err:.3e} vs scale {scale:.3e}'
    # and the split kernel against the oracle, through the shared full-size check
    test_block_full_size_vs_oracle_sample(N, d, de, nh, B)


@pytest.mark.parametrize('N,B,ks,rmp', [(256, 3, 4, 0.0), (188, 4, 2, 0.0), (512, 2, 8, 0.0), (300, 2, 4, 0.1)])
def test_fused_backward_key_split(N, B, ks, rmp, monkeypatch):
    """Same for the d_e = 8 kernels (csrc/fused_bwd.cu: fused_bwd_key_splits, forced with EGT_FUSED_KSPLIT), including a
    training step with random key masks: the counters of the mask are positions in the graph, not in the CTA's run."""
    from tests.test_parity_gpu import test_block_full_size_vs_oracle_sample, _spec_kwargs
    import egt_b200
    from oracle import egt_oracle as O
    d, de, nh = 64, 8, 8
    cfg = O.BlockConfig(model_width=d, edge_width=de, num_heads=nh, scale_degree=True, random_mask_prob=rmp)
    params = O.init_block_params(cfg, dtype=torch.float64)
    h, e, mask = O.synthetic_batch(B, N, d, de, ragged=True, dtype=torch.float32)
    blk = egt_b200.EGTBlock(**_spec_kwargs(cfg))
    blk.load_keras_weights(params)
    blk = blk.to(DEV)
    blk.train(rmp > 0)
    g = torch.Generator().manual_seed(5)
    dh = torch.randn(B, N, d, generator=g).bfloat16().to(DEV)
    de_ = torch.randn(B, N, N, de, generator=g).bfloat16().to(DEV)

    def run():
        blk.rng.offset = 41                                  # the same random mask in both runs
        hg, eg = h.bfloat16().to(DEV).requires_grad_(True), e.bfloat16().to(DEV).requires_grad_(True)
        blk.flat.grad = None
        h2, e2 = blk(hg, eg, mask.to(DEV))
        torch.autograd.backward([h2, e2], [dh, de_])
        return hg.grad.clone(), eg.grad.clone(), blk.flat.grad.clone()

    monkeypatch.setenv('EGT_FUSED_KSPLIT', '1')
    ref = run()
    monkeypatch.setenv('EGT_FUSED_KSPLIT', str(ks))
    got = run()
    assert torch.equal(got[1], ref[1])                       # de: per (row, key), untouched by the split
    for a_, b_, what in zip(got, ref, ('dh', 'de', 'weight gradients')):
        scale = float(b_.float().abs().max())
        err = float((a_.float() - b_.float()).abs().max())
        tol = 8e-3 if a_.dtype == torch.bfloat16 else 2e-3
        assert err <= tol * scale, f'{what}: {err:.3e} vs scale {scale:.3e}'
    if rmp == 0:
        test_block_full_size_vs_oracle_sample(N, d, de, nh, B)
