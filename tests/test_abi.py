"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/nvae_b200.h
declares, and the ctypes structs match the header layouts.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from nvae_tf_b200 import _lib


def test_header_parses_every_prototype():
    protos = _lib.parse_header()
    src = open(_lib.HEADER).read()
    declared = set(re.findall(r"NVAE_API[^;(]*?\b(nvae_\w+)\s*\(", src))
    assert declared == set(protos), declared ^ set(protos)
    assert len(protos) >= 35
    res, args = protos["nvae_conv2d_fwd"]
    assert res is C.c_int and len(args) == 11
    assert protos["nvae_conv2d_ws_bytes"][0] is C.c_size_t


def test_library_exports_every_declared_symbol(lib_built):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_built], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    declared = set(_lib.parse_header())
    assert declared <= exported, declared - exported
    # nothing but the ABI leaks out of the library
    assert all(s.startswith("nvae_") for s in exported if not s.startswith("_")), exported


def test_ctypes_binding_loads_without_gpu(lib_built):
    lib = _lib.lib()
    assert lib._nvae_version() == 100
    assert b"sm_100a" in lib._nvae_build_info()
    for name in _lib.parse_header():
        assert hasattr(lib, "_" + name)


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.NvaeConvDesc) == 16 * 4 + 2 * 4
    assert C.sizeof(_lib.NvaeSnLayer) == 6 * 8 + 8 * 4
    src = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER).read(), flags=re.S)
    for struct in (_lib.NvaeConvDesc, _lib.NvaeSnLayer):
        body = re.search(r"typedef struct \{([^}]*)\} " + struct.__name__ + ";", src).group(1)
        names = [n.strip() for decl in body.split(";") if decl.strip()
                 for n in decl.strip().split(" ", 1)[1].split(",")]
        assert names == [f[0] for f in struct._fields_]


def test_pure_host_entry_points(lib_built):
    """Workspace-size queries are pure functions: callable without a device."""
    lib = _lib.lib()
    d = _lib.NvaeConvDesc()
    d.N, d.H, d.W, d.Cin, d.Cout, d.R, d.S, d.stride, d.Ho, d.Wo, d.pad_t, d.pad_l = 4, 8, 8, 16, 16, 3, 3, 1, 8, 8, 1, 1
    assert lib._nvae_conv2d_ws_bytes(C.byref(d), 2) > 0
    d.Ho = 5  # inconsistent geometry is rejected (size query returns 0, launchers return NVAE_E_BADSHAPE)
    assert lib._nvae_conv2d_ws_bytes(C.byref(d), 0) == 0
    assert lib._nvae_bn_ws_bytes(1024, 64) > 0
    assert lib._nvae_se_bwd_ws_bytes(8, 64, 4) == 8 * (2 * 64 + 4) * 4


def test_launchers_reject_bad_arguments_before_touching_the_device(lib_built):
    lib = _lib.lib()
    # argument validation happens on the host, before any CUDA call, so these are safe without a GPU
    assert lib._nvae_bn_stats(None, 16, 6, None, None, None, None, 1, 0.05, 1e-5, None, None, 0, None) == -1  # C % 4
    assert lib._nvae_latent_fwd(None, None, None, 4, 16, 20, None, None, None, None, None, None) == -4
    assert lib._nvae_bernoulli_ll_fwd(None, None, 4, 32, 32, 3, 2, 0, None, None) == -1  # Cl must be C or 1
    with pytest.raises(_lib.NvaeError):
        lib.adamax(None, None, None, None, 7, None, 0.9, 0.999, 1e-7, 1.0, None)  # n % 4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NvaeError, match="no CPU fallback"):
        _lib._Lib()
