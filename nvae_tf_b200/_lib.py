"""ctypes binding of libnvae_b200.so -- the C-ABI boundary of `include/nvae_b200.h`.

The prototypes are parsed from the header itself so the Python side can never drift from the
ABI a TensorFlow custom-op wrapper (tf_op/nvae_ops.cc) binds.  There is NO fallback: if the
library is missing or a launcher returns non-zero, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "nvae_b200.h")
LIB_PATH = os.environ.get("NVAE_LIB") or os.path.join(_HERE, "libnvae_b200.so")  # NVAE_LIB: development builds

NVAE_ACT_NONE, NVAE_ACT_SWISH, NVAE_ACT_ELU = 0, 1, 2
NVAE_PREC_FP32, NVAE_PREC_TF32, NVAE_PREC_TF32X3 = 0, 1, 2
SN_ROWS_PER_CHUNK = 64

_ERRORS = {-1: "NVAE_E_BADSHAPE", -2: "NVAE_E_UNSUPPORTED", -3: "NVAE_E_WORKSPACE", -4: "NVAE_E_NULLPTR",
           -5: "NVAE_E_DRIVER"}


class NvaeError(RuntimeError):
    pass


class NvaeConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("N", "H", "W", "Cin", "Cin2", "Cout", "R", "S", "stride", "Ho", "Wo", "pad_t", "pad_l", "precision",
                 "y_ld", "y_off")] + [("pre_scale", C.c_float), ("pre_shift", C.c_float)]


class NvaeSnLayer(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("w_off", "u_off", "v_off", "t_off", "rnd_off", "tr_off")] + \
               [(n, C.c_int32) for n in ("rows", "cout", "taps", "cin", "cin_pad", "cout_pad", "chunk0", "n_chunks")]


_CTYPES = {
    "int": C.c_int, "float": C.c_float, "int64_t": C.c_int64, "int32_t": C.c_int32, "uint64_t": C.c_uint64,
    "size_t": C.c_size_t, "nvae_stream_t": C.c_void_p, "void": None,
}


def _ctype_of(decl: str):
    decl = decl.replace("const", "").strip()
    if "*" in decl:
        return C.c_char_p if decl.replace(" ", "") == "char*" else C.c_void_p
    return _CTYPES[decl.split()[0]]


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """Returns {symbol: (restype, [argtypes])} for every NVAE_API prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"NVAE_API\s+([\w\s\*]+?)\s*\b(nvae_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        if args.strip() not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                ty = a if "*" in a else a.rsplit(" ", 1)[0]
                if "*" in a:
                    ty = a[: a.rindex("*") + 1]
                argtypes.append(_ctype_of(ty))
        out[name] = (_ctype_of(ret), argtypes)
    return out


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise NvaeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C nvae_tf_b200/csrc).  nvae_tf_b200 has no CPU fallback.")
        self.dll = C.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.launches = 0  # kernels' launcher calls issued (bench.py reports gpu_launches from this)
        for name, (res, args) in self.protos.items():
            fn = getattr(self.dll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
            setattr(self, "_" + name, fn)

    def __getattr__(self, name):
        # checked call: lib.conv2d_fwd(...) -> nvae_conv2d_fwd, raises on non-zero status
        fn = self.__dict__.get("_nvae_" + name)
        if fn is None:
            raise AttributeError(name)
        if fn.restype is not C.c_int:
            return fn

        def call(*a):
            rc = fn(*a)
            self.launches += 1
            if rc != 0:
                raise NvaeError(f"nvae_{name} failed: {_ERRORS.get(rc, 'cudaError_t ' + str(rc))}")
        self.__dict__[name] = call
        return call


class NoDeviceLib:
    """Stands in for the library in a layout-only Runtime: any launch is an error (no CPU path)."""

    launches = 0

    def __getattr__(self, name):
        raise NvaeError(f"nvae_{name.lstrip('_').replace('nvae_', '')}: this Runtime was created with device='cpu' "
                        "(variable layout only); kernels need a CUDA device")


def default_precision() -> int:
    """Arithmetic mode of the convolution GEMMs unless the caller picks one (NVAE_PRECISION=fp32|tf32|tf32x3).
    Default: 3xTF32 -- tcgen05 tensor cores at fp32-level accuracy (the reference computes in fp32)."""
    env = os.environ.get("NVAE_PRECISION", "").lower()
    return {"fp32": NVAE_PREC_FP32, "tf32": NVAE_PREC_TF32, "tf32x3": NVAE_PREC_TF32X3}.get(env, NVAE_PREC_TF32X3)


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
