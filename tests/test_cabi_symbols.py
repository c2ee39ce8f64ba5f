"""The C-ABI library loads on a machine without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from usrp_nfc_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "usrp_nfc_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nfc_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = _declared()
    assert names, "no declarations found"
    assert sorted(_cabi.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    L = _cabi.lib()
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), name
    assert L.nfc_abi_version() == 7


def test_default_params_match_reference_constructors():
    p = _cabi.Params()
    _cabi.lib().nfc_default_params(ctypes.byref(p))
    # transition_sink.py:12 and background.py:17
    assert (p.samp_rate, p.lo_val, p.hi_val, p.av_window, p.max_len) == (2e6, 0.1, 1.1, 2000, 50)


def test_struct_layouts():
    """Every struct of the binding has exactly the size the library was compiled with (nfc_abi_sizeof), and the field
    offsets the header implies for the records that are read as numpy arrays."""
    L = _cabi.lib()
    sizes = {0: ctypes.sizeof(_cabi.Params), 1: _cabi.EVENT_DTYPE.itemsize, 2: _cabi.SYMBOL_DTYPE.itemsize,
             3: _cabi.FRAME_DTYPE.itemsize, 4: _cabi.FRAME_TAIL_DTYPE.itemsize, 5: ctypes.sizeof(_cabi.State),
             6: ctypes.sizeof(_cabi.Stats)}
    for which, size in sizes.items():
        assert L.nfc_abi_sizeof(which) == size, (which, L.nfc_abi_sizeof(which), size)
    assert L.nfc_abi_sizeof(99) == -1
    assert (sizes[0], sizes[1], sizes[2], sizes[3], sizes[4]) == (56, 16, 16, 24, 24)
    assert [_cabi.EVENT_DTYPE.fields[k][1] for k in ("pos", "d", "v", "type")] == [0, 8, 12, 13]
    assert [_cabi.SYMBOL_DTYPE.fields[k][1] for k in ("pos", "type", "val")] == [0, 8, 9]
    assert [_cabi.FRAME_DTYPE.fields[k][1] for k in ("pos", "bit_off", "nbits", "type")] == [0, 8, 16, 20]
    assert [_cabi.FRAME_TAIL_DTYPE.fields[k][1] for k in ("nbits", "nbytes", "byte_off", "fix_flag", "parity_ok", "crc_ok")] == [0, 4, 8, 16, 17, 18]
    assert _cabi.Params.pcm_scale.offset == 52 and _cabi.State.lastL.offset == 64 and _cabi.Stats.slicer_kernel_ms.offset == 160


def test_creation_fails_loudly_without_a_gpu():
    if _cabi.lib().nfc_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.NfcError) as e:
        _cabi.Stream(2e6)
    assert "no CPU path" in str(e.value)
