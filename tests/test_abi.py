"""CPU: the C-ABI library loads, exports every symbol include/egt_b200.h declares, and the host-side
mirror of the reference interface behaves like the reference (names, argument meaning, errors).
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import egt_b200
from egt_b200 import _lib as L
from egt_b200 import ops
from tests import philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, 'include', 'egt_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(egt_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_functions()
    assert set(names) == set(L.EXPORTS), (names, L.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.egt_abi_version() == 1


def test_struct_sizes_match_header():
    # mirrors of egt_attn_cfg_t / egt_block_cfg_t: field offsets must follow natural C alignment
    assert ctypes.sizeof(L.AttnCfg) == 96                # ... seed, offset, offset_dev (device pointer)
    assert ctypes.sizeof(L.BlockCfg) == 96 + 24
    assert ctypes.sizeof(L.BlockWeights) == 14 * 8
    assert ctypes.sizeof(L.BlockFwdIO) == 12 * 8
    assert ctypes.sizeof(L.BlockBwdIO) == 14 * 8


def test_param_layout_matches_survey_counts():
    # 3d^2+3d + d^2+d + 2d + 2d_e + 2(d_e*h+h) + h*d_e+d_e
    for d, de, h in [(64, 8, 8), (128, 32, 16), (48, 48, 8), (96, 8, 8)]:
        total, layout = ops.param_layout(ops.BlockSpec(model_width=d, edge_width=de, num_heads=h))
        assert total == 3 * d * d + 3 * d + d * d + d + 2 * d + 2 * de + 2 * (de * h + h) + h * de + de
        offs = sorted(v[0] for v in layout.values())
        assert offs[0] == 0 and len(set(offs)) == 14
    total, layout = ops.param_layout(ops.BlockSpec(model_width=64, edge_width=8, num_heads=8, edge_channel_type='none'))
    assert set(layout) == {'norm_mha_gamma', 'norm_mha_beta', 'dense_qkv_kernel', 'dense_qkv_bias',
                           'dense_mha_kernel', 'dense_mha_bias'}
    total, layout = ops.param_layout(ops.BlockSpec(model_width=64, edge_width=8, num_heads=8, edge_channel_type='bias',
                                                   gate_attention=False))
    assert 'attention_gates_kernel' not in layout and 'norm_edge_gamma' not in layout and 'dense_edge_b_kernel' in layout


def test_reference_error_behaviour():
    with pytest.raises(ValueError):                      # egt_layers.py:20-21
        egt_b200.EGT(scale_degree=True, gate_input=False)
    with pytest.raises(ValueError):                      # egt_layers.py:23-24
        egt_b200.EGT(scaler_type='cube')
    with pytest.raises(ValueError):                      # graph_xformer_model_base.py:46-47
        egt_b200.EGTBlock(model_width=16, edge_width=8, num_heads=4, gate_attention=False, scale_degree=True)
    with pytest.raises(KeyError):                        # dispatch dict graph_xformer_model_base.py:328-334
        egt_b200.EGTBlock(model_width=16, edge_width=8, num_heads=4, edge_channel_type='sparse')
    with pytest.raises(NotImplementedError):
        egt_b200.EGTBlock(model_width=16, edge_width=8, num_heads=4, add_n_norm=True)
    # no CPU fallback: the product path refuses CPU tensors loudly
    blk = egt_b200.EGTBlock(model_width=16, edge_width=8, num_heads=4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        blk(torch.zeros(1, 3, 16), torch.zeros(1, 3, 3, 8), torch.ones(1, 3, dtype=torch.bool))


def test_egt_layer_config_and_mask_contract():
    layer = egt_b200.EGT(num_heads=4, scale_degree=True, num_virtual_nodes=2, random_mask_prob=0.1)
    cfg = layer.get_config()
    for k in ('num_heads', 'clip_logits_value', 'scale_degree', 'scaler_type', 'edge_input', 'gate_input',
              'attn_mask', 'num_virtual_nodes', 'random_mask_prob'):       # egt_layers.py:42-55
        assert k in cfg
    m = torch.ones(2, 3, dtype=torch.bool)
    assert layer.compute_mask(None, [m, None, None])[0] is m              # egt_layers.py:215-217


def test_keras_weight_names_roundtrip():
    blk = egt_b200.EGTBlock(tag='03', model_width=16, edge_width=8, num_heads=4)
    w = blk.keras_weights()
    assert set(w) == {f'{n}_03/{p}' for n, p in [
        ('norm_mha', 'gamma'), ('norm_mha', 'beta'), ('dense_qkv', 'kernel'), ('dense_qkv', 'bias'),
        ('dense_mha', 'kernel'), ('dense_mha', 'bias'), ('norm_edge', 'gamma'), ('norm_edge', 'beta'),
        ('attention_gates', 'kernel'), ('attention_gates', 'bias'), ('dense_edge_b', 'kernel'),
        ('dense_edge_b', 'bias'), ('dense_edge_r', 'kernel'), ('dense_edge_r', 'bias')]}
    assert w['dense_qkv_03/kernel'].shape == (16, 48) and w['dense_edge_r_03/kernel'].shape == (4, 8)
    w2 = {k: torch.randn_like(v) for k, v in w.items()}
    blk.load_keras_weights(w2)
    for k, v in blk.keras_weights().items():
        assert torch.equal(v, w2[k])
    with pytest.raises(KeyError):
        blk.load_keras_weights({})


def test_rng_host_hook_matches_numpy_replica():
    lib = L.load()
    idx = np.array([0, 1, 2, 7, 8, 9, 1000003, 2**35 + 17], dtype=np.uint64)
    for seed, off, sid in [(0, 0, 0), (31, 9, 0), (77, 5, 1), (2**63 + 5, 2**41 + 1, 1)]:
        u = philox.uniform(seed, off, sid, idx)
        v = np.array([lib.egt_rng_uniform_host(seed, off, sid, int(i)) for i in idx], dtype=np.float32)
        assert np.array_equal(u, v)
    big = philox.noise_tensor(3, 1, 0, 4, 32, 8)
    assert 0.09 < (big < 0.1).mean() < 0.11 and big.min() > 0 and big.max() < 1


def test_rng_element_index_is_a_bijection_with_the_documented_layout():
    """include/egt_b200.h / csrc/common.cuh rng_elem_index: distinct elements draw distinct uniforms, and one
    Philox call (index >> 3) covers two consecutive keys x four consecutive heads of one query row."""
    for B, N, h in [(2, 5, 8), (1, 4, 4), (2, 7, 16), (1, 1, 8), (1, 6, 6)]:
        idx = philox.elem_index(B, N, h)
        assert idx.shape == (B, N, N, h)
        assert np.unique(idx).size == idx.size
        call, lane = idx >> np.uint64(3), idx & np.uint64(7)
        for b in range(B):
            for l in range(N):
                for m in range(N):
                    for hh in range(h):
                        same = call[b, l] == call[b, l, m, hh]
                        mm, hq = np.nonzero(same)
                        assert set(mm) <= {m - (m & 1), m - (m & 1) + 1} and set(hq // 4) == {hh // 4}
                        assert lane[b, l, m, hh] == (m & 1) * 4 + (hh & 3)


def test_workspace_query_is_monotone():
    lib = L.load()
    spec = ops.BlockSpec(model_width=64, edge_width=8, num_heads=8)
    c = spec.c_cfg(4, 32, L.EGT_BF16, False, 0, 0)
    f = lib.egt_block_workspace_bytes(ctypes.byref(c), 0)
    b = lib.egt_block_workspace_bytes(ctypes.byref(c), 1)
    assert 0 < f <= b


def test_ffn_host_side_mirrors_reference_names(golden_dir):
    """EGTFFN / EGTLayer: hidden = round(width * ffn_multiplier) (graph_xformer_model_base.py:236), flat
    parameter layout, and weight loading by the reference's layer names (no GPU compute here)."""
    from tests.golden_util import block_case
    ffn = egt_b200.EGTFFN(8, channel='edge', tag='03', ffn_multiplier=2.)
    assert (ffn.width, ffn.hidden) == (8, 16)
    assert egt_b200.EGTFFN(48, ffn_multiplier=1.5).hidden == 72
    total, layout = ops.ffn_layout(64, 128)
    assert layout['lr1_kernel'][1] == (64, 128) and layout['lr2_kernel'][1] == (128, 64)
    assert all(off % 4 == 0 for off, _ in layout.values()) and total >= 2 * 64 * 128 + 128 + 3 * 64
    w = {f'{stem}_edge_03/{wn}': torch.full(shape, float(i))
         for i, ((stem, wn), shape) in enumerate([(('norm_fnn', 'gamma'), (8,)), (('norm_fnn', 'beta'), (8,)),
                                                  (('fnn_lr1', 'kernel'), (8, 16)), (('fnn_lr1', 'bias'), (16,)),
                                                  (('fnn_lr2', 'kernel'), (16, 8)), (('fnn_lr2', 'bias'), (8,))])}
    ffn.load_keras_weights(w)
    assert float(ffn.view('lr2_kernel')[3, 5]) == 4.0 and float(ffn.view('norm_beta')[0]) == 1.0
    with pytest.raises(KeyError):
        egt_b200.EGTFFN(8, channel='node', tag='03').load_keras_weights(w)
    c = block_case(golden_dir, 0)
    layer = egt_b200.EGTLayer(model_width=c['cfg'].model_width, edge_width=c['cfg'].edge_width,
                              num_heads=c['cfg'].num_heads)
    layer.load_keras_weights(c['params'])
    assert layer.ffn_edge is not None and layer.ffn_node.width == c['cfg'].model_width
    with pytest.raises(RuntimeError):                      # no CPU fallback
        layer.ffn_node(torch.zeros(2, 3, c['cfg'].model_width))


def test_ffn_workspace_and_allreduce_sizes_are_host_functions():
    """egt_ffn_workspace_bytes / egt_peer_allreduce_push_floats are pure host arithmetic (callable without a GPU): which
    feed-forward shapes ask for a workspace (the cuBLAS path of csrc/node_blas.cu) and how large the symmetric buffer of
    the push all-reduce must be."""
    lib = L.load()

    def ws(rows, width, hidden, act=L.EGT_ACT_ELU, dtype=L.EGT_BF16):
        c = L.FfnCfg()
        c.rows, c.width, c.hidden, c.dtype, c.activation, c.ln_eps = rows, width, hidden, dtype, act, 1e-3
        return int(lib.egt_ffn_workspace_bytes(ctypes.byref(c)))

    assert ws(128 * 128 * 128, 8, 16) == 0                   # tcgen05 kernels: no workspace
    assert ws(128 * 128, 64, 128) == 0
    assert ws(32 * 512, 128, 256) > 32 * 512 * 256 * 2       # node channel at d = 128: cuBLAS path
    assert ws(64 * 190, 96, 192) > 0
    assert ws(64 * 190, 64, 96) > 0                          # hidden != 2 w: not a tcgen05 shape
    assert ws(100, 128, 256) == 0                            # too few rows for a GEMM to pay: CUDA-core kernels
    assert ws(32 * 512, 128, 256, act=L.EGT_ACT_RELU) == 0   # relu stays on the float32 kernels
    assert ws(32 * 512, 128, 256, dtype=L.EGT_F32) == 0
    assert int(lib.egt_peer_allreduce_push_floats(17000, 8)) == 4 * 17000 * 8
