"""Reference-side binding of libegt_b200.so -- the file a maintainer of shamim-hussain/egt would add as
``lib/models/egt_b200_binding.py`` to run ``EGT.call_gated / call_ungated`` (lib/models/egt_layers.py:57-143,145-213)
on the B200 kernels.  ctypes only: no build step on the reference side.

    from lib.models.egt_b200_binding import egt_attention
    # inside EGT.call_gated, instead of the einsum / softmax / sigmoid chain:
    V_att, H_hat = egt_attention(QKV, E, G, mask[0], num_heads=self.num_heads, clip=self.clip_logits_value,
                                 scale_degree=self.scale_degree, scaler_type=self.scaler_type,
                                 num_virtual_nodes=self.num_virtual_nodes)

TensorFlow is imported lazily (it is absent from the image this repository is developed in); everything that does
not need it -- the structure layouts, the DLPack capsule -> device pointer helper, the error mapping -- is exercised
by tests/test_integration_stub.py against the built library and PyTorch's DLPack capsules.
"""
import ctypes as C
import os

EGT_F32, EGT_BF16 = 0, 1
EGT_SCALER = {'log': 0, 'linear': 1}


class AttnCfg(C.Structure):            # mirrors egt_attn_cfg_t (include/egt_b200.h)
    _fields_ = [('B', C.c_int32), ('N', C.c_int32), ('h', C.c_int32), ('dk', C.c_int32),
                ('dtype', C.c_int32), ('edge_input', C.c_int32), ('gate_input', C.c_int32),
                ('attn_mask', C.c_int32), ('has_clip', C.c_int32),
                ('clip_lo', C.c_float), ('clip_hi', C.c_float),
                ('scale_degree', C.c_int32), ('scaler_type', C.c_int32),
                ('num_virtual_nodes', C.c_int32), ('training', C.c_int32),
                ('random_mask_prob', C.c_float), ('attn_dropout', C.c_float),
                ('seed', C.c_uint64), ('offset', C.c_uint64), ('offset_dev', C.c_void_p)]


# ---- DLPack (dlpack.h v0.x as TensorFlow and PyTorch export it) -----------------------------------------
class _DLDevice(C.Structure):
    _fields_ = [('device_type', C.c_int32), ('device_id', C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [('code', C.c_uint8), ('bits', C.c_uint8), ('lanes', C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [('data', C.c_void_p), ('device', _DLDevice), ('ndim', C.c_int32), ('dtype', _DLDataType),
                ('shape', C.POINTER(C.c_int64)), ('strides', C.POINTER(C.c_int64)), ('byte_offset', C.c_uint64)]


class _DLManagedTensor(C.Structure):
    _fields_ = [('dl_tensor', _DLTensor), ('manager_ctx', C.c_void_p), ('deleter', C.c_void_p)]


_get_pointer = C.pythonapi.PyCapsule_GetPointer
_get_pointer.restype = C.c_void_p
_get_pointer.argtypes = [C.py_object, C.c_char_p]


def device_ptr(capsule):
    """Address of the first element of the tensor a DLPack capsule ('dltensor') describes.  The capsule must stay
    alive (and un-consumed) for as long as the pointer is used."""
    mt = C.cast(_get_pointer(capsule, b'dltensor'), C.POINTER(_DLManagedTensor)).contents
    return (mt.dl_tensor.data or 0) + mt.dl_tensor.byte_offset


_lib = None


def load(path=None):
    """dlopen libegt_b200.so (EGT_B200_LIB, or next to the egt_b200 package) and declare the prototypes used here."""
    global _lib
    if _lib is None:
        path = path or os.environ.get('EGT_B200_LIB') or os.path.join(
            os.path.dirname(os.path.abspath(__file__)), '..', 'egt_b200', 'lib', 'libegt_b200.so')
        lib = C.CDLL(path)
        lib.egt_last_error.restype = C.c_char_p
        vp = C.c_void_p
        lib.egt_attn_fwd.argtypes = [C.POINTER(AttnCfg), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.egt_attn_bwd.argtypes = [C.POINTER(AttnCfg), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        _lib = lib
    return _lib


def check(rc):
    """egt_status -> the exception class the reference raises for the same mistake (egt_layers.py:20-24,70)."""
    if rc == 0:
        return
    msg = load().egt_last_error().decode()
    if rc == -1:
        raise AssertionError(msg)          # EGT_E_SHAPE  <-> assert at egt_layers.py:70
    if rc == -6:
        raise ValueError(msg)              # EGT_E_ARG    <-> ValueError at egt_layers.py:20-24
    raise RuntimeError(msg)


def make_cfg(B, N, h, dk, *, bf16=False, edge_input=True, gate_input=True, clip=(-5., 5.), scale_degree=False,
             scaler_type='log', num_virtual_nodes=0, training=False, random_mask_prob=0., attn_dropout=0., seed=0,
             offset=0):
    c = AttnCfg()
    c.B, c.N, c.h, c.dk = B, N, h, dk
    c.dtype = EGT_BF16 if bf16 else EGT_F32
    c.edge_input, c.gate_input, c.attn_mask = int(edge_input), int(gate_input), 0
    c.has_clip = int(clip is not None)
    c.clip_lo, c.clip_hi = (clip if clip is not None else (0., 0.))
    c.scale_degree, c.scaler_type, c.num_virtual_nodes = int(scale_degree), EGT_SCALER[scaler_type], num_virtual_nodes
    c.training, c.random_mask_prob, c.attn_dropout = int(training), random_mask_prob, attn_dropout
    c.seed, c.offset = seed, offset
    return c


def egt_attention(QKV, E, G, mask, *, num_heads, clip=(-5., 5.), scale_degree=False, scaler_type='log',
                  num_virtual_nodes=0, stream=None):
    """TensorFlow entry point: ``(QKV [B,N,3d], E [B,N,N,h], G [B,N,N,h], mask [B,N] bool) -> (V_att, H_hat)`` with the
    gradient wired through ``tf.custom_gradient``.  float32 tensors on one GPU (the reference's dtype)."""
    import tensorflow as tf                      # noqa: deferred -- see the module docstring
    to_dl = tf.experimental.dlpack.to_dlpack
    lib = load()
    B, N = int(QKV.shape[0]), int(QKV.shape[1])
    d = int(QKV.shape[2]) // 3
    cfg = make_cfg(B, N, num_heads, d // num_heads, clip=clip, scale_degree=scale_degree, scaler_type=scaler_type,
                   num_virtual_nodes=num_virtual_nodes)
    st = C.c_void_p(stream or 0)                 # the caller's stream handle; 0 = the legacy default stream

    @tf.custom_gradient
    def op(qkv, e, g):
        m8 = tf.cast(mask, tf.uint8)
        v_att = tf.zeros([B, N, d], tf.float32)
        h_hat = tf.zeros_like(e)
        lse = tf.zeros([2, B, N, num_heads], tf.float32)
        deg = tf.zeros([B, N, num_heads], tf.float32)
        caps = [to_dl(t) for t in (qkv, e, g, m8, v_att, h_hat, lse, deg)]
        p = [C.c_void_p(device_ptr(c_)) for c_ in caps]
        check(lib.egt_attn_fwd(C.byref(cfg), p[0], p[1], p[2], None, p[3], p[4], p[5], None, p[6], p[7], st))

        def grad(d_v_att, d_h_hat):
            d_qkv, dE, dG = tf.zeros_like(qkv), tf.zeros_like(e), tf.zeros_like(g)
            row_ws = tf.zeros([2, B, N, num_heads], tf.float32)
            gc = [to_dl(t) for t in (qkv, e, g, m8, lse, deg, d_v_att, d_h_hat, d_qkv, dE, dG, row_ws)]
            q = [C.c_void_p(device_ptr(c_)) for c_ in gc]
            check(lib.egt_attn_bwd(C.byref(cfg), q[0], q[1], q[2], None, q[3], q[4], q[5], q[6], q[7], q[8], q[9], q[10],
                                   q[11], st))
            return d_qkv, dE, dG

        return (v_att, h_hat), grad

    return op(QKV, E, G)
