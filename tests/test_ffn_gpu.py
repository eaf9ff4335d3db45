"""GPU parity of the feed-forward half of a layer (reference: ffnlr1 / ffnact / ffnlr2,
lib/models/graph_xformer_model_base.py:229-258, chained by ffn_block :309-324) through the C ABI
(egt_ffn_fwd / egt_ffn_bwd) against the golden vectors made from the reference's source and the oracle."""
import pytest
import torch

from oracle import egt_oracle as O
from tests.golden_util import block_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = {torch.float32: dict(rtol=1e-3, atol=2e-4), torch.bfloat16: dict(rtol=1e-2, atol=1e-2)}


def _close(got, ref, dtype, what):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    tol = TOL[dtype]
    bound = tol['atol'] * max(1.0, float(ref.abs().max())) + tol['rtol'] * ref.abs()
    err = (got - ref).abs()
    assert not (err > bound).any(), f'{what}: max err {float(err.max()):.3e}'


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('idx', [0, 1, 2, 3, 5, 8, 11, 13])
def test_ffn_forward_golden(idx, dtype, golden_dir):
    """h_ffn / e_ffn of the golden block cases: FFN applied to the attention block's outputs."""
    import egt_b200
    c = block_case(golden_dir, idx)
    cfg = c['cfg']
    if cfg.add_n_norm:
        pytest.skip('add_n_norm=True (post-norm) is not built')
    for channel, x, ref, width in (('node', c['h_out'], c['h_ffn'], cfg.model_width),
                                   ('edge', c['e_out'], c.get('e_ffn'), cfg.edge_width)):
        if ref is None or f'ffn_{channel}/lr1/kernel' not in c['params']:
            continue
        ffn = egt_b200.EGTFFN(width, channel=channel, ffn_multiplier=cfg.ffn_multiplier, activation=cfg.activation)
        ffn.load_keras_weights(c['params'])
        ffn = ffn.to(DEV)
        y = ffn(x.to(dtype).to(DEV))
        yr = O.ffn_channel(x.to(dtype).double(), c['params'], f'ffn_{channel}', cfg)
        _close(y, yr, dtype, f'case {idx} {channel} ffn vs oracle')
        if dtype == torch.float32:
            _close(y, ref, dtype, f'case {idx} {channel} ffn vs golden')


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('rows,width,mult,act', [((3, 17), 64, 2.0, 'elu'), ((2, 9, 9), 8, 2.0, 'elu'),
                                                 ((2, 7, 7), 32, 2.0, 'elu'), ((5, 11), 48, 2.0, 'relu'),
                                                 ((2, 5), 96, 2.0, 'elu'), ((2, 6, 6), 64, 1.5, 'tanh'),
                                                 ((1, 300), 128, 2.0, 'elu')])
def test_ffn_forward_backward_vs_oracle(rows, width, mult, act, dtype):
    import egt_b200
    torch.manual_seed(width + len(rows))
    ffn = egt_b200.EGTFFN(width, ffn_multiplier=mult, activation=act)
    with torch.no_grad():
        ffn.flat.add_(0.1 * torch.randn_like(ffn.flat))
    params = {f'ffn_node/{n}/{w}': ffn.view(f).double().clone().requires_grad_(True)
              for f, (n, w) in {'norm_gamma': ('norm', 'gamma'), 'norm_beta': ('norm', 'beta'), 'lr1_kernel': ('lr1', 'kernel'),
                                'lr1_bias': ('lr1', 'bias'), 'lr2_kernel': ('lr2', 'kernel'), 'lr2_bias': ('lr2', 'bias')}.items()}
    ffn = ffn.to(DEV)
    x = torch.randn(*rows, width)
    dy = torch.randn(*rows, width)
    xg = x.to(dtype).to(DEV).requires_grad_(True)
    y = ffn(xg)
    gx, gflat = torch.autograd.grad(y, [xg, ffn.flat], dy.to(dtype).to(DEV))
    cfg = O.BlockConfig(model_width=width, edge_width=8, num_heads=8, ffn_multiplier=mult, activation=act)
    xr = x.to(dtype).double().requires_grad_(True)
    yr = O.ffn_channel(xr, params, 'ffn_node', cfg)
    gr = torch.autograd.grad(yr, [xr] + list(params.values()), dy.to(dtype).double())
    _close(y, yr, dtype, 'y')
    _close(gx, gr[0], dtype, 'dx')
    ffn.flat.grad = gflat
    tol = 1e-3 if dtype == torch.float32 else 1e-2
    for (name, _), ref in zip(params.items(), gr[1:]):
        field = {'norm/gamma': 'norm_gamma', 'norm/beta': 'norm_beta', 'lr1/kernel': 'lr1_kernel', 'lr1/bias': 'lr1_bias',
                 'lr2/kernel': 'lr2_kernel', 'lr2/bias': 'lr2_bias'}[name.split('/', 1)[1]]
        got = ffn.grad_view(field).double().cpu()
        err = float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-6)
        assert err < tol, f'grad {name}: rel-to-max err {err:.3e}'


@pytest.mark.parametrize('idx', [0, 2, 5])
def test_full_layer_golden(idx, golden_dir):
    """edge_update + ffn_block of one layer (graph_xformer_model_base.py:335-341) against the golden h_ffn / e_ffn."""
    import egt_b200
    from tests.test_parity_gpu import _spec_kwargs
    c = block_case(golden_dir, idx)
    cfg = c['cfg']
    if cfg.add_n_norm or c['training']:
        pytest.skip('post-norm / random-mask cases are covered elsewhere')
    layer = egt_b200.EGTLayer(ffn_multiplier=cfg.ffn_multiplier, activation=cfg.activation, **_spec_kwargs(cfg))
    layer.load_keras_weights(c['params'])
    layer = layer.to(DEV)
    em = c.get('edge_mask')
    h2, e2 = layer(c['h'].float().to(DEV), c['e'].float().to(DEV), c['mask'].to(DEV),
                   edge_mask=None if em is None else em.to(DEV), training=False)
    _close(h2, c['h_ffn'], torch.float32, f'layer case {idx} h')
    if 'e_ffn' in c and layer.ffn_edge is not None:
        _close(e2, c['e_ffn'], torch.float32, f'layer case {idx} e')
