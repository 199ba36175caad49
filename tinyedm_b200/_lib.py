"""ctypes binding of libtinyedm_b200.so (the C ABI declared in include/tinyedm_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised. PyTorch only provides device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtinyedm_b200.so")

_lib = None
_initialised_devices: set[int] = set()

_P = c_void_p
_I = c_int
_F = c_float

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_SIGNATURES = {
    "tedm_version": [],
    "tedm_init": [_I],
    "tedm_conv2d_forward": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P, _F, _P, _I, _F, c_uint64, _I, _P],
    "tedm_conv2d_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P],
}
_RESTYPES = {"tedm_last_error": c_char_p}


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"tinyedm_b200: native library not found at {LIB_PATH}; build it with `make` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        lib.tedm_last_error.restype = c_char_p
        lib.tedm_last_error.argtypes = []
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        _lib = lib
    return _lib


def exported_symbols() -> list[str]:
    return ["tedm_last_error", *_SIGNATURES.keys()]


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.tedm_last_error()
        raise RuntimeError(f"{name} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


def init_device(index: int) -> None:
    if index not in _initialised_devices:
        call("tedm_init", index)
        _initialised_devices.add(index)
