"""ctypes binding of include/yolov3_b200.h.  There is NO fallback: if the shared library is missing, or a call
fails, this raises — the product path never routes around the CUDA extension."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libyolov3_b200.so"
_lib = None


class Y3Error(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """struct y3_conv_desc (include/yolov3_b200.h)."""

    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("c_in", C.c_int32), ("c_out", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("act", C.c_int32),
        ("in_", C.c_void_p), ("in_ld", C.c_int32), ("in_coff", C.c_int32),
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("out", C.c_void_p), ("out_ld", C.c_int32), ("out_coff", C.c_int32),
        ("res", C.c_void_p), ("res_ld", C.c_int32), ("res_coff", C.c_int32),
        ("upsample", C.c_int32),
        ("raw", C.c_void_p), ("na", C.c_int32), ("no", C.c_int32),
        ("err", C.c_void_p),
    ]


def _declare(lib):
    i32, vp, sz = C.c_int32, C.c_void_p, C.c_size_t
    sigs = {
        "y3_version": ([], C.c_int),
        "y3_last_error": ([C.c_char_p, sz], C.c_int),
        "y3_device_check": ([], C.c_int),
        "y3_conv_bn_act_fwd": ([C.POINTER(ConvDesc), vp], C.c_int),
        "y3_conv_cout_pad": ([i32], C.c_int),
        "y3_conv_first_fwd": ([vp, i32, i32, i32, vp, vp, i32, vp, i32, i32, vp], C.c_int),
        "y3_nchw_to_padded_nhwc": ([vp, i32, i32, i32, i32, vp, i32, i32, vp], C.c_int),
        "y3_padded_nhwc_to_nchw": ([vp, i32, i32, i32, i32, i32, i32, vp, vp], C.c_int),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    return sigs


SYMBOLS: dict = {}


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise Y3Error(
                f"{_LIB_PATH} is missing: build it with `python -m yolov3_b200.build` (or __graft_entry__.build()). "
                "yolov3_b200 has no CPU or PyTorch fallback path."
            )
        _lib = C.CDLL(str(_LIB_PATH))
        SYMBOLS.update(_declare(_lib))
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().y3_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise Y3Error(f"{what or 'yolov3_b200'} failed (rc={rc}): {last_error()}")
