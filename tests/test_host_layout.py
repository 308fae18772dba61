"""CPU tests of the host mirror in layout-only mode (device='cpu': variables, names, tables -- no launches)."""
import numpy as np
import pytest
import torch

import helpers as H
from nvae_tf_b200 import _lib, parallel
from nvae_tf_b200.models import NVAE
from nvae_tf_b200.runtime import same_pad, conv_desc, Runtime
from oracle import nvae_oracle as O


@pytest.fixture(scope="module")
def default_model(lib_built):
    cfg = O.NVAEConfig()
    return cfg, NVAE(**H.mirror_kwargs(cfg, 144), device="cpu")


def test_default_model_matches_reference_structure(default_model):
    cfg, m = default_model
    assert m.count_params() == 40_128_893
    assert len(m.rt.sn_convs) == 163 and len(m.rt.bn_loss_layers) == 88
    params, trainable, bnl, _ = O.build_params(cfg)
    assert set(m.rt.variables) == set(params)
    for k, v in m.rt.variables.items():
        assert v.shape == tuple(params[k].shape), k
        assert v.trainable == (k in set(trainable)), k
    assert sorted(b.gamma.name[: -len("/gamma")] for b in m.rt.bn_loss_layers) == sorted(bnl)
    np.testing.assert_allclose(m.calculate_kl_alphas(2, [5, 10]), O.kl_alphas(cfg))


def test_arena_layout_is_aligned_and_disjoint(default_model):
    _, m = default_model
    spans = sorted((v.offset, v.offset + v.size) for v in m.rt.trainable_variables)
    assert all(lo % 4 == 0 for lo, _ in spans)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert m.rt.params.numel() % 4 == 0 and m.rt.grads.shape == m.rt.params.shape


def test_sn_table_covers_every_wrapped_conv(default_model):
    _, m = default_model
    tab = m.rt.sn_host
    assert len(tab) == 163
    chunks = 0
    for i, sn in enumerate(m.rt.sn_convs):
        k = sn.layer.kernel
        assert tab[i].w_off == k.offset and tab[i].u_off == sn.u.offset
        assert tab[i].rows * tab[i].cout == k.size and tab[i].taps * tab[i].cin == tab[i].rows
        assert tab[i].chunk0 == chunks
        chunks += tab[i].n_chunks
    assert chunks == m.rt.sn_chunk_layer.numel()


def test_layout_only_runtime_cannot_launch(default_model):
    _, m = default_model
    with pytest.raises(_lib.NvaeError):
        m.train_step(np.zeros((2, 32, 32, 1), np.float32))


def test_same_pad_and_conv_desc():
    assert same_pad(32, 3, 2) == (16, 0) and same_pad(8, 5, 1) == (8, 2) and same_pad(31, 1, 2) == (16, 0)
    rt = Runtime(device="cpu")
    d = conv_desc(rt, (4, 32, 32, 8), 0, (1, 1, 8, 4), 2, shift=(1, 1), y_ld=16, y_off=4)
    assert (d.Ho, d.Wo, d.pad_t, d.pad_l, d.y_ld, d.y_off) == (16, 16, -1, -1, 16, 4)
    with pytest.raises(ValueError):
        conv_desc(rt, (4, 8, 8, 8), 4, (1, 1, 8, 4), 1)


def test_bucket_bounds_cover_the_arena_back_to_front():
    b = parallel.bucket_bounds(10, 4)
    assert [(s.start, s.stop) for s in b] == [(6, 10), (2, 6), (0, 2)]
    assert parallel.bucket_bounds(10, 0) == [slice(0, 10)]
    assert parallel.shard_batch(8, 1, 2) == slice(4, 8)
    with pytest.raises(ValueError):
        parallel.shard_batch(9, 0, 2)


def test_bench_traffic_record_is_tied_to_the_kernel_source(tmp_path, monkeypatch):
    """bench.py's roofline.traffic comes from profiles/dominant_kernel_traffic.json and is reported only while the hash of
    conv_tc.cu recorded with the ncu capture still matches the source (no stale literal)."""
    import hashlib
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rec = json.load(open(os.path.join(root, "profiles", "dominant_kernel_traffic.json")))
    sha = hashlib.sha256(open(os.path.join(root, "nvae_tf_b200", "csrc", "conv_tc.cu"), "rb").read()).hexdigest()[:16]
    assert rec["f16x2"]["conv_tc_sha16"] == sha, "conv_tc.cu changed since the ncu capture: re-capture or accept traffic = null"
    val, src = bench.dominant_traffic(True, True, True)
    assert val == rec["f16x2"]["dram_bytes"] and "r02_ncu_conv" in src
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))  # no profiles/ there -> no number, and it says why
    assert bench.dominant_traffic(True, True, True)[0] is None
