"""Loader of the CPU logic emulator build (TEST INFRASTRUCTURE ONLY; see cuda_emu.h)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_PATH = os.path.join(_HERE, "libdeltaq_emu.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return EMU_PATH


_lib = None


def library():
    global _lib
    if _lib is None:
        from deltaq_b200._native import Library
        build()
        _lib = Library(EMU_PATH)
    return _lib


def context():
    from deltaq_b200._native import Context
    return Context(lib=library())
