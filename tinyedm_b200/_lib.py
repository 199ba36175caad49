"""ctypes binding of libtinyedm_b200.so (the C ABI declared in include/tinyedm_b200.h).

The signatures are parsed from the header itself, so the header is the single source of truth for the
boundary. There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised. PyTorch only provides device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtinyedm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "tinyedm_b200.h")

_lib = None
_fns: dict[str, object] = {}
_initialised_devices: set[int] = set()


class WeightDesc(ctypes.Structure):
    """Mirror of `tedm_weight_desc` (include/tinyedm_b200.h)."""

    _fields_ = [
        ("w", c_void_p), ("grad", c_void_p), ("g_hat", c_void_p), ("out_fwd", c_void_p), ("out_dgrad", c_void_p),
        ("out_f32", c_void_p), ("stats", c_void_p),
        ("rows", ctypes.c_int32), ("cin", ctypes.c_int32), ("taps", ctypes.c_int32), ("kpad", ctypes.c_int32),
        ("row_start", ctypes.c_int32), ("qkv_head_dim", ctypes.c_int32), ("group_start", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


def _ctype_of(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return c_void_p
    base = decl.rsplit(" ", 1)[0].replace("const", "").strip()
    return {"int": c_int, "float": c_float, "uint64_t": c_uint64, "int64_t": c_int64, "tedm_stream_t": c_void_p}[base]


def header_signatures() -> dict[str, list]:
    """Parses `int tedm_xxx(...)` / `const char* tedm_xxx(void)` prototypes out of the public header."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    sigs: dict[str, list] = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(tedm_\w+)\s*\(([^)]*)\)\s*;", text):
        args = m.group(3).strip()
        sigs[m.group(2)] = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
    return sigs


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"tinyedm_b200: native library not found at {LIB_PATH}; build it with `make` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in header_signatures().items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.argtypes = argtypes
            fn.restype = c_char_p if name == "tedm_last_error" else c_int
            _fns[name] = fn
        _lib = lib
    return _lib


def exported_symbols() -> list[str]:
    return list(header_signatures().keys())


_n_calls = 0


def n_calls() -> int:
    """C-ABI calls issued so far by this process (== kernel launches of this library, to first order)."""
    return _n_calls


def call(name: str, *args) -> None:
    global _n_calls
    if _lib is None:
        load()
    _n_calls += 1
    rc = _fns[name](*args)
    if rc != 0:
        msg = _fns["tedm_last_error"]()
        raise RuntimeError(f"{name} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


def call_int(name: str, *args) -> int:
    """For the few entry points whose return value is a quantity rather than a status (version, chunk size)."""
    if _lib is None:
        load()
    return int(_fns[name](*args))


def init_device(index: int) -> None:
    if index not in _initialised_devices:
        call("tedm_init", index)
        _initialised_devices.add(index)
