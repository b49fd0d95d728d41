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


# ---- a plain C caller (tests/c_abi_caller.c): the header is valid C99 and the library needs no Python ------------------

def _build_c_caller(tmp_path, libdir, libname):
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / ("c_caller_" + libname))
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_caller.c"), "-L", libdir, "-l" + libname,
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_c_caller_without_a_device(tmp_path):
    """dq_cuda_create says DQ_ERR_NO_DEVICE (no CPU fallback); the host-only entry points work from C."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from deltaq_b200 import build
    build.build()
    exe = _build_c_caller(tmp_path, os.path.join(ROOT, "deltaq_b200"), "deltaq_cuda")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no device" in r.stdout and "host-only entry points ok" in r.stdout


def test_c_caller_device_path_on_the_emulator(tmp_path):
    """The same program linked against the CPU logic emulator build (test infrastructure): its device checks -- Sort,
    Diff.Create's streams, Search + the host consumer, the patch file and Patch.Apply -- as a C caller makes them."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu
    emu.build()
    exe = _build_c_caller(tmp_path, os.path.dirname(emu.EMU_PATH), "deltaq_emu")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "device path ok" in r.stdout


@pytest.mark.gpu
def test_c_caller_on_the_gpu(tmp_path):
    import subprocess
    exe = _build_c_caller(tmp_path, os.path.join(ROOT, "deltaq_b200"), "deltaq_cuda")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "device path ok" in r.stdout
