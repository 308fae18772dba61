"""CPU tests of the tcgen05 convolution launch planner (nvae_conv2d_plan_info: a pure host function): which
arithmetic each of the model's convolutions gets, that the tile / ring / TMEM budgets hold for every convolution of
the default-config step at batch 144, and that the switches documented in include/nvae_b200.h do what they say."""
import ctypes as C

import pytest

from nvae_tf_b200 import _lib

KEYS = ["ok", "BN", "m_tiles", "n_tiles", "KU", "ctas", "stages", "a_slots", "acc_bufs", "f16", "nsub", "dual", "split",
        "smem", "ws_kb", "chunks"]


def desc(N, H, W, Cin, Cout, k, stride=1, Cin2=0, precision=None):
    d = _lib.NvaeConvDesc()
    d.N, d.H, d.W, d.Cin, d.Cin2, d.Cout, d.R, d.S, d.stride = N, H, W, Cin, Cin2, Cout, k, k, stride
    d.Ho, d.Wo = -(-H // stride), -(-W // stride)
    tot_h = max((d.Ho - 1) * stride + k - H, 0)
    tot_w = max((d.Wo - 1) * stride + k - W, 0)
    d.pad_t, d.pad_l = tot_h // 2, tot_w // 2
    d.precision = _lib.NVAE_PREC_TF32X3 if precision is None else precision
    return d


def plan(lib, d, which):
    out = (C.c_int32 * 16)()
    ok = lib._nvae_conv2d_plan_info(C.byref(d), which, out)
    p = dict(zip(KEYS, list(out)))
    assert p["ok"] == ok
    return p


def tmem_columns(p):
    """Accumulator + A-ring columns the kernel uses (conv_tc.cu: tmem_a_ring, a_slot)."""
    if not p["f16"]:
        return p["acc_bufs"] * p["BN"] + p["a_slots"] * 64
    if p["dual"]:
        return 2 * p["BN"] + p["a_slots"] * 64
    return p["acc_bufs"] * p["BN"] + p["a_slots"] * 32


@pytest.fixture
def lib(lib_built, monkeypatch):
    for k in ("NVAE_F16X3", "NVAE_F16X3_MIN_GFLOP", "NVAE_F16X3_DUAL", "NVAE_F16X3_NSUB", "NVAE_WGRAD_CHUNKS", "NVAE_TC_PAIR"):
        monkeypatch.delenv(k, raising=False)
    return _lib.lib()


def test_large_gemms_run_3xfp16_with_shared_operand_tiles(lib):
    # postprocess 5x5 384 -> 384 at 16x16, batch 144: N = 384 is one tile of two accumulators sharing each A tile
    for which in (0, 1, 2):
        p = plan(lib, desc(144, 16, 16, 384, 384, 5), which)
        assert p["ok"] and p["f16"] == 1 and p["nsub"] == 2 and p["dual"] == 0 and p["BN"] == 384 and p["n_tiles"] == 1
        assert p["acc_bufs"] == 1 and tmem_columns(p) <= 512 and p["smem"] <= 227 * 1024
    # 5x5 192 -> 192 at 32x32: two M tiles per CTA share each B tile
    for which in (0, 1, 2):
        p = plan(lib, desc(144, 32, 32, 192, 192, 5), which)
        assert p["ok"] and p["f16"] == 1 and p["dual"] == 1 and p["nsub"] == 1 and p["BN"] == 192
        assert tmem_columns(p) <= 512 and p["smem"] <= 227 * 1024
    # its filter gradient reads 226 MB with 19 dual tiles: the pixel-aligned 7-way K split, not equal unit ranges
    p = plan(lib, desc(144, 32, 32, 192, 192, 5), 2)
    assert p["m_tiles"] == 38 and p["ctas"] == 19 * 7 and p["split"] == 1
    # the 384 layer's operands (113 MB) fit L2: equal unit ranges over all 148 SMs
    assert plan(lib, desc(144, 16, 16, 384, 384, 5), 2)["ctas"] == 148


def test_small_and_ineligible_convolutions_stay_3xtf32(lib):
    for d in (desc(144, 4, 4, 256, 256, 3), desc(144, 8, 8, 128, 768, 1), desc(144, 32, 32, 32, 32, 3),
              desc(144, 8, 8, 128, 256, 3, stride=2)):
        for which in (0, 1, 2):
            p = plan(lib, d, which)
            assert p["ok"] and p["f16"] == 0 and p["dual"] == 0 and p["nsub"] == 1, (which, p)
            assert tmem_columns(p) <= 512 and p["smem"] <= 227 * 1024
    # two sources (DecoderSampleCombiner concat) are never 3xFP16, whatever the size
    assert plan(lib, desc(4096, 16, 16, 384, 384, 5, Cin2=32), 0)["f16"] == 0
    # fp32 mode and shapes the tensor-core path does not take report no plan
    assert plan(lib, desc(144, 16, 16, 384, 384, 5, precision=_lib.NVAE_PREC_FP32), 0)["ok"] == 0
    assert plan(lib, desc(144, 32, 32, 1, 32, 3), 0)["ok"] == 0  # Cin = 1 stem


def test_switches(lib, monkeypatch):
    big = desc(144, 16, 16, 384, 384, 5)
    monkeypatch.setenv("NVAE_F16X3", "0")
    p = plan(lib, big, 0)
    assert p["f16"] == 0 and p["BN"] == 192 and p["n_tiles"] == 2
    monkeypatch.delenv("NVAE_F16X3")
    monkeypatch.setenv("NVAE_F16X3_NSUB", "0")
    p = plan(lib, big, 0)
    assert p["f16"] == 1 and p["nsub"] == 1 and p["BN"] == 192 and p["dual"] == 1
    monkeypatch.setenv("NVAE_F16X3_DUAL", "0")
    p = plan(lib, big, 0)
    assert p["f16"] == 1 and p["nsub"] == 1 and p["dual"] == 0
    monkeypatch.delenv("NVAE_F16X3_NSUB")
    monkeypatch.delenv("NVAE_F16X3_DUAL")
    small = desc(8, 4, 4, 64, 64, 3)
    assert plan(lib, small, 0)["f16"] == 0
    monkeypatch.setenv("NVAE_F16X3_MIN_GFLOP", "0")
    assert plan(lib, small, 0)["f16"] == 1
    monkeypatch.delenv("NVAE_F16X3_MIN_GFLOP")
    assert plan(lib, big, 2)["chunks"] == 1
    monkeypatch.setenv("NVAE_WGRAD_CHUNKS", "4")
    assert plan(lib, big, 2)["chunks"] == 4
    monkeypatch.setenv("NVAE_WGRAD_CHUNKS", "5")  # 144 % 5 != 0 -> the next count that divides the batch
    assert plan(lib, big, 2)["chunks"] == 4


def test_workspace_covers_partials_and_packed_operands(lib):
    big = desc(144, 16, 16, 384, 384, 5)
    w_bytes = 5 * 5 * 384 * 384 * 4
    dy_bytes = 144 * 16 * 16 * 384 * 4
    for which, extra in ((0, w_bytes), (1, w_bytes), (2, dy_bytes)):
        p = plan(lib, big, which)
        partial = p["ctas"] * 2 * 128 * p["BN"] * 4 if p["split"] else 0
        assert p["ws_kb"] * 1024 >= partial + extra, (which, p)
        assert lib._nvae_conv2d_ws_bytes(C.byref(big), which) >= p["ws_kb"] * 1024
