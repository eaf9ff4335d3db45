"""ctypes binding of libegt_b200.so (the C ABI declared in include/egt_b200.h).

The product path fails loudly when the CUDA library is missing: there is no CPU or PyTorch
fallback anywhere in this package.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libegt_b200.so')

EGT_F32, EGT_BF16 = 0, 1
EGT_SCALER_LOG, EGT_SCALER_LINEAR = 0, 1
EGT_EDGE_NONE, EGT_EDGE_BIAS, EGT_EDGE_RESIDUAL, EGT_EDGE_CONSTRAINED = 0, 1, 2, 3
EGT_ACT_NONE, EGT_ACT_LRELU, EGT_ACT_RELU, EGT_ACT_ELU, EGT_ACT_TANH, EGT_ACT_SIGMOID = range(6)
EGT_MASK_NONE, EGT_MASK_DENSE, EGT_MASK_ADJ_U8 = 0, 1, 2
EGT_E_SHAPE, EGT_E_DTYPE, EGT_E_ALIGN, EGT_E_ARCH, EGT_E_CUDA, EGT_E_ARG = -1, -2, -3, -4, -5, -6

EXPORTS = ['egt_abi_version', 'egt_last_error', 'egt_last_path', 'egt_rng_uniform_host',
           'egt_block_param_layout', 'egt_attn_fwd', 'egt_attn_bwd', 'egt_block_workspace_bytes',
           'egt_block_fwd', 'egt_block_bwd', 'egt_launch_count', 'egt_profile_enable', 'egt_profile_read',
           'egt_debug_umma_probe', 'egt_debug_mma_timing', 'egt_debug_force_staged', 'egt_peer_allreduce', 'egt_peer_allreduce_push', 'egt_peer_allreduce_push_floats',
           'egt_ffn_fwd', 'egt_ffn_bwd', 'egt_ffn_workspace_bytes', 'egt_ffn_fwd_ws', 'egt_ffn_bwd_ws']

FFN_FIELDS = ['norm_gamma', 'norm_beta', 'lr1_kernel', 'lr1_bias', 'lr2_kernel', 'lr2_bias']


class FfnCfg(C.Structure):
    """egt_ffn_cfg_t (include/egt_b200.h)."""
    _fields_ = [('rows', C.c_int64), ('width', C.c_int32), ('hidden', C.c_int32), ('dtype', C.c_int32),
                ('activation', C.c_int32), ('ln_eps', C.c_float)]


class FfnWeights(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in FFN_FIELDS]


class FfnGrads(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in FFN_FIELDS]


WEIGHT_FIELDS = ['norm_mha_gamma', 'norm_mha_beta', 'dense_qkv_kernel', 'dense_qkv_bias',
                 'dense_mha_kernel', 'dense_mha_bias', 'norm_edge_gamma', 'norm_edge_beta',
                 'attention_gates_kernel', 'attention_gates_bias', 'dense_edge_b_kernel',
                 'dense_edge_b_bias', 'dense_edge_r_kernel', 'dense_edge_r_bias']


class AttnCfg(C.Structure):
    _fields_ = [('B', C.c_int32), ('N', C.c_int32), ('h', C.c_int32), ('dk', C.c_int32),
                ('dtype', C.c_int32), ('edge_input', C.c_int32), ('gate_input', C.c_int32),
                ('attn_mask', C.c_int32), ('has_clip', C.c_int32), ('clip_lo', C.c_float),
                ('clip_hi', C.c_float), ('scale_degree', C.c_int32), ('scaler_type', C.c_int32),
                ('num_virtual_nodes', C.c_int32), ('training', C.c_int32),
                ('random_mask_prob', C.c_float), ('attn_dropout', C.c_float),
                ('seed', C.c_uint64), ('offset', C.c_uint64), ('offset_dev', C.c_void_p)]


class BlockCfg(C.Structure):
    _fields_ = [('attn', AttnCfg), ('d_e', C.c_int32), ('edge_channel_type', C.c_int32),
                ('gate_attention', C.c_int32), ('edge_act', C.c_int32),
                ('edge_act_alpha', C.c_float), ('ln_eps', C.c_float)]


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in WEIGHT_FIELDS]


class BlockGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in WEIGHT_FIELDS]


class BlockFwdIO(C.Structure):
    _fields_ = [('h', C.c_void_p), ('e', C.c_void_p), ('mask', C.c_void_p), ('adj', C.c_void_p),
                ('h_out', C.c_void_p), ('e_out', C.c_void_p), ('qkv', C.c_void_p),
                ('v_att', C.c_void_p), ('lse', C.c_void_p), ('deg', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


class BlockBwdIO(C.Structure):
    _fields_ = [('h', C.c_void_p), ('e', C.c_void_p), ('mask', C.c_void_p), ('adj', C.c_void_p),
                ('qkv', C.c_void_p), ('v_att', C.c_void_p), ('lse', C.c_void_p), ('deg', C.c_void_p),
                ('dh_out', C.c_void_p), ('de_out', C.c_void_p), ('dh', C.c_void_p), ('de', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


_lib = None


def load():
    """Load libegt_b200.so (built in-tree by `python -m egt_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -m egt_b200.build` '
            '(nvcc, sm_100a). egt_b200 has no CPU / PyTorch fallback.')
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.egt_abi_version.restype = C.c_int
    lib.egt_last_error.restype = C.c_char_p
    lib.egt_last_path.restype = C.c_int
    lib.egt_rng_uniform_host.restype = C.c_float
    lib.egt_rng_uniform_host.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
    lib.egt_block_param_layout.restype = C.c_int64
    lib.egt_block_param_layout.argtypes = [C.POINTER(BlockCfg), C.POINTER(C.c_int64)]
    lib.egt_attn_fwd.restype = C.c_int
    lib.egt_attn_fwd.argtypes = [C.POINTER(AttnCfg)] + [vp] * 11
    lib.egt_attn_bwd.restype = C.c_int
    lib.egt_attn_bwd.argtypes = [C.POINTER(AttnCfg)] + [vp] * 14
    lib.egt_block_workspace_bytes.restype = C.c_size_t
    lib.egt_block_workspace_bytes.argtypes = [C.POINTER(BlockCfg), C.c_int32]
    lib.egt_block_fwd.restype = C.c_int
    lib.egt_block_fwd.argtypes = [C.POINTER(BlockCfg), C.POINTER(BlockWeights), C.POINTER(BlockFwdIO), vp]
    lib.egt_block_bwd.restype = C.c_int
    lib.egt_block_bwd.argtypes = [C.POINTER(BlockCfg), C.POINTER(BlockWeights), C.POINTER(BlockGrads),
                                  C.POINTER(BlockBwdIO), vp]
    lib.egt_ffn_fwd.restype = C.c_int
    lib.egt_ffn_fwd.argtypes = [C.POINTER(FfnCfg), C.POINTER(FfnWeights), vp, vp, vp]
    lib.egt_ffn_bwd.restype = C.c_int
    lib.egt_ffn_bwd.argtypes = [C.POINTER(FfnCfg), C.POINTER(FfnWeights), C.POINTER(FfnGrads), vp, vp, vp, vp]
    lib.egt_ffn_workspace_bytes.restype = C.c_size_t
    lib.egt_ffn_workspace_bytes.argtypes = [C.POINTER(FfnCfg)]
    lib.egt_ffn_fwd_ws.restype = C.c_int
    lib.egt_ffn_fwd_ws.argtypes = [C.POINTER(FfnCfg), C.POINTER(FfnWeights), vp, vp, vp, C.c_size_t, vp]
    lib.egt_ffn_bwd_ws.restype = C.c_int
    lib.egt_ffn_bwd_ws.argtypes = [C.POINTER(FfnCfg), C.POINTER(FfnWeights), C.POINTER(FfnGrads), vp, vp, vp, vp, C.c_size_t, vp]
    lib.egt_peer_allreduce.restype = C.c_int
    lib.egt_peer_allreduce.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, vp]
    lib.egt_peer_allreduce_push.restype = C.c_int
    lib.egt_peer_allreduce_push.argtypes = [vp, vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, vp]
    lib.egt_peer_allreduce_push_floats.restype = C.c_int64
    lib.egt_peer_allreduce_push_floats.argtypes = [C.c_int64, C.c_int]
    lib.egt_launch_count.restype = C.c_long
    lib.egt_profile_enable.argtypes = [C.c_int]
    lib.egt_profile_read.restype = C.c_int
    lib.egt_profile_read.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int]
    lib.egt_debug_force_staged.argtypes = [C.c_int]
    lib.egt_debug_umma_probe.restype = C.c_int
    lib.egt_debug_umma_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp, vp, vp, vp]
    if lib.egt_abi_version() != 1:
        raise RuntimeError('libegt_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc):
    """Map a negative egt_status onto the exception class the reference raises
    (ValueError for bad flags egt_layers.py:20-24, AssertionError for bad channel counts :70)."""
    if rc == 0:
        return
    msg = load().egt_last_error().decode()
    if rc == EGT_E_ARG:
        raise ValueError(msg)
    if rc == EGT_E_SHAPE:
        raise AssertionError(msg)
    if rc == EGT_E_DTYPE:
        raise TypeError(msg)
    raise RuntimeError(msg)


def profile_read(max_entries=32):
    """{kernel name: (total ms, launches)} since egt_profile_enable(1)."""
    lib = load()
    names = C.create_string_buffer(64 * max_entries)
    ms = (C.c_double * max_entries)()
    cnt = (C.c_long * max_entries)()
    n = lib.egt_profile_read(names, ms, cnt, max_entries)
    return {names.raw[64 * i:64 * (i + 1)].split(b'\0')[0].decode(): (ms[i], cnt[i]) for i in range(n)}
