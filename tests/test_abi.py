"""The C-ABI library loads and exports every symbol include/deltaq_cuda.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "deltaq_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dq_cuda_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    from deltaq_b200 import _native
    assert sorted(_native.EXPORTS) == declared_symbols()


def test_cuda_library_builds_and_exports_every_symbol():
    from deltaq_b200 import build
    try:
        path = build.build()
    except RuntimeError as e:
        pytest.skip(str(e))
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from deltaq_b200 import CudaSuffixSort, _native
    with pytest.raises(_native.NativeError) as ei:
        CudaSuffixSort()
    assert ei.value.status == _native.DQ_ERR_NO_DEVICE
