"""ctypes binding of libnk_b200.so (the C ABI declared in include/nk_b200.h).

The signatures are parsed from the header itself so the binding cannot drift from the ABI.
There is no fallback: if the shared library is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
HEADER = _PKG.parent / "include" / "nk_b200.h"
LIB_PATH = _PKG / "csrc" / "libnk_b200.so"


class NkError(RuntimeError):
    pass


class nk_operand(ctypes.Structure):
    _fields_ = [
        ("ptr", ctypes.c_void_p),
        ("mn_major", ctypes.c_int32),
        ("conv", ctypes.c_int32),
        ("inner", ctypes.c_int64),
        ("rows", ctypes.c_int64),
        ("row_stride", ctypes.c_int64),
        ("nb2", ctypes.c_int64),
        ("b2_stride", ctypes.c_int64),
        ("nb1", ctypes.c_int64),
        ("b1_stride", ctypes.c_int64),
        ("H", ctypes.c_int32),
        ("W", ctypes.c_int32),
        ("nimg", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
    ]


class nk_gemm_desc(ctypes.Structure):
    _fields_ = [
        ("A", nk_operand),
        ("B", nk_operand),
        ("M", ctypes.c_int32),
        ("N", ctypes.c_int32),
        ("K", ctypes.c_int32),
        ("nb2", ctypes.c_int32),
        ("nb1", ctypes.c_int32),
        ("ksize", ctypes.c_int32),
        ("pad", ctypes.c_int32),
        ("wgrad", ctypes.c_int32),
        ("C", ctypes.c_void_p),
        ("ldc", ctypes.c_int64),
        ("c_b2_stride", ctypes.c_int64),
        ("c_b1_stride", ctypes.c_int64),
        ("out", ctypes.c_int32),
        ("epi", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("rows_per_img", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("bias_img", ctypes.c_void_p),
        ("residual", ctypes.c_void_p),
        ("ldr", ctypes.c_int64),
        ("rowvec", ctypes.c_void_p),
        ("aux", ctypes.c_void_p),
        ("force_bn", ctypes.c_int32),
        ("force_splits", ctypes.c_int32),
        ("force_cta_group", ctypes.c_int32),
        ("force_dual", ctypes.c_int32),
    ]


_CTYPE = {
    "int": ctypes.c_int,
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "nk_stream_t": ctypes.c_void_p,
}


def _map_type(t: str):
    t = t.replace("const", "").strip()
    if t.endswith("*"):
        base = t[:-1].strip()
        if base == "nk_gemm_desc":
            return ctypes.POINTER(nk_gemm_desc)
        if base == "char":
            return ctypes.c_char_p
        return ctypes.c_void_p
    return _CTYPE[t]


def parse_header(path: Path = HEADER) -> dict[str, tuple]:
    """{function name: (restype, [argtypes])} for every `nk_*` prototype in the header."""
    text = re.sub(r"/\*.*?\*/", "", path.read_text(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int64_t|int|const char\s*\*)\s+(nk_\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)  # strip the parameter name
                argtypes.append(_map_type(mm.group(1).strip()))
        protos[name] = (_map_type(ret) if "char" in ret else _CTYPE[ret.strip()], argtypes)
    return protos


def _load() -> ctypes.CDLL:
    path = Path(os.environ.get("NK_B200_LIB", LIB_PATH))
    if "NK_B200_LIB" not in os.environ:
        # a fresh clone has no binary (the .so is git-ignored) and an edited source tree has a stale one: (re)build in
        # tree when nvcc is available — the same routine `__graft_entry__.build()` runs
        from . import build as _build
        stale = (not path.exists()) or _build._stale(path, _build._sources() + _build._headers())
        if stale and Path(_build.NVCC).exists() and not os.environ.get("NK_B200_NO_AUTOBUILD"):
            _build.build(verbose=True)
    if not path.exists():
        raise ImportError(
            f"neurosis_b200: kernel library {path} is missing - run `python -m neurosis_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)"
        )
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()
PROTOTYPES = parse_header()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.nk_last_error()
        raise NkError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
