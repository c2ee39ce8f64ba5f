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
    assert L.nfc_abi_version() == 3


def test_default_params_match_reference_constructors():
    p = _cabi.Params()
    _cabi.lib().nfc_default_params(ctypes.byref(p))
    # transition_sink.py:12 and background.py:17
    assert (p.samp_rate, p.lo_val, p.hi_val, p.av_window, p.max_len) == (2e6, 0.1, 1.1, 2000, 50)


def test_struct_layouts():
    assert ctypes.sizeof(_cabi.Params) == 64 or ctypes.sizeof(_cabi.Params) % 8 == 0
    assert _cabi.EVENT_DTYPE.itemsize == 16 and _cabi.SYMBOL_DTYPE.itemsize == 16 and _cabi.FRAME_DTYPE.itemsize == 24


def test_creation_fails_loudly_without_a_gpu():
    if _cabi.lib().nfc_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.NfcError) as e:
        _cabi.Stream(2e6)
    assert "no CPU path" in str(e.value)
