"""The C-ABI shared library loads and exports every symbol include/air_b200.h declares;
without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes
import re

import pytest
import torch

import air_b200 as ab


def _declared_in_header():
    src = open(ab._cabi.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(air_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_in_header()
    assert len(names) >= 10
    raw = ctypes.CDLL(ab._cabi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/air_b200.h but not exported"
    assert sorted(ab._cabi.declared_symbols()) == names, "ctypes signature table out of sync with the header"
    assert ab._cabi.lib().air_abi_version() >= 1


def test_no_cpu_fallback():
    with pytest.raises(ab.AirError):
        ab.transformer(torch.zeros(1, 4, 4, 1), torch.zeros(1, 6), (2, 2))
    if not torch.cuda.is_available():
        a = ctypes.c_int()
        assert ab._cabi.lib().air_device_info(a, a, a) < 0
        assert b"no CPU fallback" in ab._cabi.lib().air_last_error()


def test_bad_arguments_are_rejected_before_launch():
    l = ab._cabi.lib()
    assert l.air_st_forward(None, None, None, 1, 4, 4, 1, 2, 2, None) == -5   # AIR_ERR_NULL
    assert l.air_st_forward(None, None, None, 0, 4, 4, 1, 2, 2, None) == 0    # empty batch: no pointers needed
    one = ctypes.c_void_p(16)
    assert l.air_st_forward(one, one, one, 1, 0, 4, 1, 2, 2, None) == -1       # AIR_ERR_BAD_SHAPE
    assert b"bad shape" in l.air_last_error()
    assert l.air_st_forward(one, one, one, 0, 4, 4, 1, 2, 2, None) == 0        # empty batch is a no-op
