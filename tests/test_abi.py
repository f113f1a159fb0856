"""CPU: the C-ABI library loads, exports every symbol include/mms_b200.h declares, and rejects bad
arguments with the documented error convention — no compute call is made (no GPU here)."""
import ctypes

import pytest

from multimodalstudio_b200 import _lib


def test_header_parses_and_every_symbol_is_exported():
    protos = _lib.parse_header()
    assert len(protos) >= 36
    lib = _lib.load_library()
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/mms_b200.h but not exported"
    assert lib.mmsb_version().decode().startswith("mms_b200")
    assert isinstance(_lib.launch_count(), int)


def test_error_convention_maps_to_python_exceptions():
    lib = _lib.load_library()
    d = _lib.MmsbHashGridDesc()
    d.num_levels, d.features_per_level, d.log2_hashmap_size = 99, 2, 19          # too many levels
    rc = lib.mmsb_hashgrid_fwd(ctypes.byref(d), None, 3, None, None, None, 32, None, 10, None)
    assert rc == -1 and "num_levels" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "hashgrid_fwd")
    d.num_levels, d.features_per_level = 16, 3                                     # unsupported F
    d.log2_hashmap_size = 19
    rc = lib.mmsb_hashgrid_fwd(ctypes.byref(d), ctypes.c_void_p(8), 3, ctypes.c_void_p(8), None, ctypes.c_void_p(8), 64, None, 10, None)
    assert rc == -2
    assert lib.mmsb_linear_fwd(None, 2, None, None, None, 4, 10, 4, 4, 0, 1.0, None) == -1        # ldx < in_dim
    assert lib.mmsb_neus_weights_fwd(None, None, None, None, None, None, 1.0, None, 2000, 10, None) == -1   # s > 1024
    # empty inputs are a no-op success without touching the device
    assert lib.mmsb_linear_fwd(None, 4, None, None, None, 4, 0, 4, 4, 0, 1.0, None) == 0
    assert lib.mmsb_composite_fwd(None, None, None, 3, None, None, None, None, None, None, None, 8, 0, None) == 0


def test_cpu_tensors_are_rejected_not_computed():
    import torch
    from multimodalstudio_b200 import ops
    desc = ops.make_hashgrid_desc(16, 2, 12, ops.hash_resolutions(16, 1024, 16))
    with pytest.raises(ValueError):
        ops.HashGridFn.apply(torch.rand(4, 3), torch.rand(16 * 4096, 2), None, desc)
    with pytest.raises(ValueError):
        ops.mlp_forward(torch.rand(4, 8), [torch.rand(8, 8)], [None], "ReLU", "None")
