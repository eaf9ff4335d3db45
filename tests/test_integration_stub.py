"""The reference-side binding INTEGRATION.md shows (integration/egt_b200_binding.py) is real code: its structure
layout matches the library's, its DLPack helper returns the tensor's address, its status mapping raises the
reference's exception classes.  (The TensorFlow half needs TensorFlow, which this image does not have.)"""
import ctypes as C
import importlib.util
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _binding():
    spec = importlib.util.spec_from_file_location('egt_b200_binding', os.path.join(ROOT, 'integration', 'egt_b200_binding.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_struct_layout_matches_the_python_host_side():
    b = _binding()
    from egt_b200 import _lib as L
    assert C.sizeof(b.AttnCfg) == C.sizeof(L.AttnCfg)
    assert [f[0] for f in b.AttnCfg._fields_] == [f[0] for f in L.AttnCfg._fields_]


def test_dlpack_capsule_gives_the_tensor_address():
    b = _binding()
    t = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    cap = torch.utils.dlpack.to_dlpack(t)
    assert b.device_ptr(cap) == t.data_ptr()
    v = t[1:]                                             # a view: DLPack carries the offset in data / byte_offset
    assert b.device_ptr(torch.utils.dlpack.to_dlpack(v)) == v.data_ptr()


def test_library_loads_and_status_maps_to_reference_exceptions():
    b = _binding()
    lib = b.load(os.path.join(ROOT, 'egt_b200', 'lib', 'libegt_b200.so'))
    cfg = b.make_cfg(1, 4, 8, 2, scale_degree=True, gate_input=False)      # egt_layers.py:20-21: needs gates
    rc = lib.egt_attn_fwd(C.byref(cfg), None, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0
    with pytest.raises(ValueError):
        b.check(rc)
