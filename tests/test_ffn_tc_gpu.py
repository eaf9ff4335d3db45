"""GPU parity of the tensor-core feed-forward kernels (csrc/ffn_tc.cu; reference: ffnlr1 / ffnact / ffnlr2 and
ffn_block, lib/models/graph_xformer_model_base.py:229-258, :309-324) through egt_ffn_fwd / egt_ffn_bwd: against the
oracle, against the CUDA-core kernels of the same library, at sizes with several tiles per CTA and a ragged last
tile, and through the linearity of the backward in dy."""
import os

import pytest
import torch

from oracle import egt_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
FIELDS = {'norm_gamma': ('norm', 'gamma'), 'norm_beta': ('norm', 'beta'), 'lr1_kernel': ('lr1', 'kernel'),
          'lr1_bias': ('lr1', 'bias'), 'lr2_kernel': ('lr2', 'kernel'), 'lr2_bias': ('lr2', 'bias')}


def _kernels_of(fn):
    """Names of this library's kernels launched by fn()."""
    from egt_b200 import _lib as L
    lib = L.load()
    lib.egt_profile_enable(1)
    out = fn()
    torch.cuda.synchronize()
    prof = L.profile_read()
    lib.egt_profile_enable(0)
    return out, set(prof)


def _make(width, act, seed, scale=0.1):
    import egt_b200
    torch.manual_seed(seed)
    ffn = egt_b200.EGTFFN(width, ffn_multiplier=2.0, activation=act)
    with torch.no_grad():
        ffn.flat.add_(scale * torch.randn_like(ffn.flat))
    params = {f'ffn_node/{n}/{w}': ffn.view(f).double().clone().requires_grad_(True) for f, (n, w) in FIELDS.items()}
    return ffn.to(DEV), params


def _run(ffn, x, dy):
    xg = x.to(DEV).requires_grad_(True)
    y = ffn(xg)
    gx, gflat = torch.autograd.grad(y, [xg, ffn.flat], dy.to(DEV))
    return y, gx, gflat


def _rel(got, ref):
    got, ref = got.detach(), ref.detach()
    return float((got.double().cpu() - ref.double().cpu()).abs().max()) / max(float(ref.abs().max()), 1e-6)


# rows chosen so that rows * width / 64 super-rows give: one ragged tile; exactly one tile; 2 tiles per CTA + a ragged one
@pytest.mark.parametrize('width,rows', [(8, 8 * 77), (8, 8 * 128), (8, 8 * (148 * 128 * 2 + 77)), (16, 4 * (148 * 128 + 5)),
                                        (32, 2 * (3 * 128 + 19)), (32, 2 * (148 * 128 * 3 + 1)), (64, 51), (64, 16384),
                                        (64, 148 * 128 + 130)])
def test_ffn_tc_vs_oracle(width, rows):
    ffn, params = _make(width, 'elu', seed=width + rows % 97)
    x = torch.randn(rows, width).bfloat16()
    dy = torch.randn(rows, width).bfloat16()
    (y, gx, gflat), names = _kernels_of(lambda: _run(ffn, x, dy))
    assert {'ffn_tc_fwd_kernel', 'ffn_tc_bwd_kernel'} <= names, names
    cfg = O.BlockConfig(model_width=width, edge_width=8, num_heads=8, ffn_multiplier=2.0, activation='elu')
    xr = x.double().requires_grad_(True)
    yr = O.ffn_channel(xr, params, 'ffn_node', cfg)
    gr = torch.autograd.grad(yr, [xr] + list(params.values()), dy.double())
    assert _rel(y, yr) < 1e-2, f'y {_rel(y, yr):.3e}'
    assert _rel(gx, gr[0]) < 1e-2, f'dx {_rel(gx, gr[0]):.3e}'
    ffn.flat.grad = gflat
    for (f, _), ref in zip(FIELDS.items(), gr[1:]):
        err = _rel(ffn.grad_view(f), ref)
        assert err < 1e-2, f'grad {f}: rel-to-max err {err:.3e}'


@pytest.mark.parametrize('act', ['linear', 'tanh', 'sigmoid'])
@pytest.mark.parametrize('width', [8, 64])
def test_ffn_tc_activations(width, act):
    """Run-time activation code of the tensor-core kernels (the compile-time one is elu)."""
    rows = 5 * 128 * 64 // width + 64 // width * 3
    ffn, params = _make(width, act, seed=7)
    x = torch.randn(rows, width).bfloat16()
    dy = torch.randn(rows, width).bfloat16()
    (y, gx, gflat), names = _kernels_of(lambda: _run(ffn, x, dy))
    assert 'ffn_tc_bwd_kernel' in names
    cfg = O.BlockConfig(model_width=width, edge_width=8, num_heads=8, ffn_multiplier=2.0, activation=act)
    xr = x.double().requires_grad_(True)
    yr = O.ffn_channel(xr, params, 'ffn_node', cfg)
    gr = torch.autograd.grad(yr, [xr] + list(params.values()), dy.double())
    assert _rel(y, yr) < 1e-2 and _rel(gx, gr[0]) < 1e-2
    ffn.flat.grad = gflat
    for (f, _), ref in zip(FIELDS.items(), gr[1:]):
        assert _rel(ffn.grad_view(f), ref) < 1e-2, f


@pytest.mark.parametrize('width', [8, 32, 64])
def test_ffn_tc_matches_cuda_core_path(width, monkeypatch):
    """Same call, both implementations of the library (EGT_FFN_TC=0 selects ffn_kernels.cu)."""
    rows = 64 // width * (9 * 128 + 31)
    ffn, _ = _make(width, 'elu', seed=3)
    x = torch.randn(rows, width).bfloat16()
    dy = torch.randn(rows, width).bfloat16()
    (y1, gx1, gf1), n1 = _kernels_of(lambda: _run(ffn, x, dy))
    monkeypatch.setenv('EGT_FFN_TC', '0')
    (y0, gx0, gf0), n0 = _kernels_of(lambda: _run(ffn, x, dy))
    assert 'ffn_tc_fwd_kernel' in n1 and 'ffn_tc_fwd_kernel' not in n0
    assert _rel(y1, y0) < 1e-2 and _rel(gx1, gx0) < 1e-2 and _rel(gf1, gf0) < 1e-2


def test_ffn_tc_backward_is_linear_in_dy():
    """Full edge tensor of the headline configuration (128 graphs x 128 x 128 x 8): bwd(a dy1 + dy2) = a bwd(dy1) + bwd(dy2)."""
    width, rows = 8, 128 * 128 * 128
    ffn, _ = _make(width, 'elu', seed=11)
    x = torch.randn(rows, width, device=DEV).bfloat16()
    dy1 = torch.randn(rows, width, device=DEV).bfloat16()
    dy2 = torch.randn(rows, width, device=DEV).bfloat16()
    dys = (2.0 * dy1.float() + dy2.float()).bfloat16()
    _, g1, f1 = _run(ffn, x, dy1)
    _, g2, f2 = _run(ffn, x, dy2)
    _, gs, fs = _run(ffn, x, dys)
    assert _rel(gs, 2.0 * g1.float() + g2.float()) < 2e-2
    assert _rel(fs, 2.0 * f1 + f2) < 1e-2
    # rows are independent: the first half of the tensor alone gives the first half of y and dx
    y_full = ffn(x)
    y_half = ffn(x[:rows // 2])
    assert torch.equal(y_full[:rows // 2], y_half)


def test_ffn_relu_stays_on_fp32_kernels():
    """relu' is discontinuous: with bf16 operands the sign of a pre-activation near zero may differ from the
    reference's, so the tensor-core path declines it and the call still matches the oracle."""
    width, rows = 8, 8 * 300
    ffn, params = _make(width, 'relu', seed=5)
    x = torch.randn(rows, width).bfloat16()
    dy = torch.randn(rows, width).bfloat16()
    (y, gx, gflat), names = _kernels_of(lambda: _run(ffn, x, dy))
    assert 'ffn_tc_fwd_kernel' not in names and 'ffn_tc_bwd_kernel' not in names
    cfg = O.BlockConfig(model_width=width, edge_width=8, num_heads=8, ffn_multiplier=2.0, activation='relu')
    xr = x.double().requires_grad_(True)
    yr = O.ffn_channel(xr, params, 'ffn_node', cfg)
    gr = torch.autograd.grad(yr, [xr], dy.double())
    assert _rel(y, yr) < 1e-2 and _rel(gx, gr[0]) < 1e-2


@pytest.mark.parametrize('width,rows,mult,act', [(96, 700, 2.0, 'elu'), (128, 1500, 2.0, 'elu'), (64, 600, 1.5, 'tanh'),
                                                 (48, 1024, 2.0, 'elu'), (128, 32 * 512, 2.0, 'elu')])
def test_ffn_blas_vs_oracle(width, rows, mult, act):
    """Channel shapes the tcgen05 kernels do not serve (node channel at d = 96 / 128, hidden != 2 w) run as cuBLAS GEMMs
    (csrc/node_blas.cu: ffn_blas_fwd / ffn_blas_bwd) in a workspace the host side allocates."""
    import egt_b200
    torch.manual_seed(rows)
    ffn = egt_b200.EGTFFN(width, ffn_multiplier=mult, activation=act)
    with torch.no_grad():
        ffn.flat.add_(0.1 * torch.randn_like(ffn.flat))
    params = {f'ffn_node/{n}/{w}': ffn.view(f).double().clone().requires_grad_(True) for f, (n, w) in FIELDS.items()}
    ffn = ffn.to(DEV)
    x = torch.randn(rows, width).bfloat16()
    dy = torch.randn(rows, width).bfloat16()
    (y, gx, gflat), names = _kernels_of(lambda: _run(ffn, x, dy))
    assert 'cublas_gemm' in names and 'ffn_fwd_kernel' not in names and 'ffn_tc_fwd_kernel' not in names, names
    cfg = O.BlockConfig(model_width=width, edge_width=8, num_heads=8, ffn_multiplier=mult, activation=act)
    xr = x.double().requires_grad_(True)
    yr = O.ffn_channel(xr, params, 'ffn_node', cfg)
    gr = torch.autograd.grad(yr, [xr] + list(params.values()), dy.double())
    assert _rel(y, yr) < 1e-2, f'y {_rel(y, yr):.3e}'
    assert _rel(gx, gr[0]) < 1e-2, f'dx {_rel(gx, gr[0]):.3e}'
    ffn.flat.grad = gflat
    for (f, _), ref in zip(FIELDS.items(), gr[1:]):
        err = _rel(ffn.grad_view(f), ref)
        assert err < 1e-2, f'grad {f}: rel-to-max err {err:.3e}'
