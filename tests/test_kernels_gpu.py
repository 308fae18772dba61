"""GPU parity tests, one kernel family at a time, through the C-ABI (ctypes) against the float64
oracle restatement (oracle/nvae_oracle.py).  fp32 CUDA-core arithmetic: tolerance 2e-5 relative
to the tensor's max magnitude (the north-star bound is 1e-3)."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import nvae_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture()
def rt(lib_built):
    from nvae_tf_b200.runtime import Runtime
    r = Runtime(seed=7)
    with r:
        yield r


def dev(rt, a, needs_grad=True):
    from nvae_tf_b200.runtime import DeviceTensor
    return DeviceTensor(torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(rt.device), needs_grad)


def seed_grad(rt, y, dy):
    y.grad = torch.as_tensor(np.ascontiguousarray(dy, dtype=np.float32)).to(rt.device)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def npy(t):
    return t.detach().cpu().numpy().astype(np.float64)


def check(name, got, want, tol=TOL, floor=0.0):
    err = H.max_rel_err(got, want, floor)
    assert err <= tol, f"{name}: max rel err {err:.3e} > {tol:.1e}"


ACTS = {0: lambda t: t, 1: O.swish, 2: O.elu}


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("training,upsample,shape", [(True, False, (6, 8, 8, 32)), (True, True, (3, 4, 4, 64)),
                                                     (False, False, (5, 7, 7, 16)), (True, False, (2, 3, 5, 1536)),
                                                     # cluster-fused path at model shapes (8 CTAs per channel chunk,
                                                     # ragged last rank), and a tensor above its size limit (split kernels)
                                                     (True, False, (144, 4, 4, 256)), (True, True, (37, 8, 8, 128)),
                                                     (True, False, (72, 32, 32, 160))])
def test_bn_act(rt, act, training, upsample, shape):
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import BatchNormalization
    rng = np.random.default_rng(0)
    C = shape[-1]
    bn = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=C, name="bn")
    rt.finalize()
    vals = {"gamma": rng.normal(1, 0.2, C), "beta": rng.normal(0, 0.3, C), "moving_mean": rng.normal(0, 0.5, C),
            "moving_variance": rng.uniform(0.5, 2, C)}
    for k, v in vals.items():
        getattr(bn, k).assign(v)
    x = f32(rng.normal(0.5, 2.0, shape))
    oshape = (shape[0], shape[1] * 2, shape[2] * 2, C) if upsample else shape
    dy = f32(rng.normal(0, 1, oshape))
    xt = dev(rt, x)
    with rt.gradient_tape() as tape:
        y = R.bn_act(rt, xt, bn, act, training, upsample=upsample)
    seed_grad(rt, y, dy)
    rt.backward(tape)
    # oracle
    p = {"bn/" + k: H.t64(f32(v)) for k, v in vals.items()}
    p["bn/gamma"].requires_grad_(True)
    p["bn/beta"].requires_grad_(True)
    c = O.Ctx(p, training)
    xo = H.t64(x).requires_grad_(True)
    yo = ACTS[act](O.batch_norm(c, "bn", xo))
    if upsample:
        yo = O.upsample_nearest2(yo)
    yo.backward(H.t64(dy))
    check("y", npy(y.data), yo.detach().numpy())
    check("dx", npy(xt.grad), xo.grad.numpy())
    check("dgamma", npy(bn.gamma.grad), p["bn/gamma"].grad.numpy())
    check("dbeta", npy(bn.beta.grad), p["bn/beta"].grad.numpy())
    if training:
        check("moving_mean", npy(bn.moving_mean.value), c.new_stats["bn/moving_mean"].numpy())
        check("moving_variance", npy(bn.moving_variance.value), c.new_stats["bn/moving_variance"].numpy())


CONV_CASES = [
    # (N, H, W, Cin, Cin2, Cout, k, stride, residual, shift)
    (4, 8, 8, 16, 0, 16, 3, 1, False, (0, 0)),      # encoder cell conv (encoder.py:92-98)
    (3, 4, 4, 32, 0, 192, 1, 1, False, (0, 0)),     # decoder cell expand (decoder.py:126-128)
    (3, 4, 4, 32, 20, 32, 1, 1, False, (0, 0)),     # DecoderSampleCombiner concat(x,z) (decoder.py:114-117)
    (3, 4, 4, 32, 0, 32, 1, 1, True, (0, 0)),       # EncoderDecoderCombiner residual (encoder.py:14-16)
    (2, 8, 8, 16, 0, 32, 3, 2, False, (0, 0)),      # Rescaler DOWN, asymmetric SAME pad (common.py:155-163)
    (2, 8, 8, 8, 0, 4, 1, 2, False, (1, 1)),        # SkipScaler shifted view (preprocess.py:69)
    (2, 8, 8, 8, 0, 4, 1, 2, False, (0, 1)),
    (2, 9, 7, 5, 0, 3, 5, 1, False, (0, 0)),        # 5x5, ragged channel counts
    (2, 32, 32, 1, 0, 8, 3, 1, False, (0, 0)),      # stem Cin=1
    (2, 16, 16, 8, 0, 1, 3, 1, False, (0, 0)),      # head Cout=1
    (1, 1, 1, 3, 0, 2, 3, 1, False, (0, 0)),        # degenerate spatial size
    (2, 6, 6, 24, 0, 40, 3, 1, False, (0, 0)),      # sampler conv, N=40
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_dgrad_wgrad(rt, case):
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import Conv2D
    N, Hh, W, Cin, Cin2, Cout, k, stride, residual, shift = case
    rng = np.random.default_rng(1)
    conv = Conv2D(Cout, (k, k), strides=(stride, stride), padding="same", in_channels=Cin + Cin2, name="c")
    rt.finalize()
    w = f32(rng.normal(0, 0.3, (k, k, Cin + Cin2, Cout)))
    b = f32(rng.normal(0, 0.3, Cout))
    conv.kernel.assign(w)
    conv.bias.assign(b)
    x = f32(rng.normal(0, 1, (N, Hh, W, Cin)))
    x2 = f32(rng.normal(0, 1, (N, Hh, W, Cin2))) if Cin2 else None
    xo, x2o = H.t64(x).requires_grad_(True), (H.t64(x2).requires_grad_(True) if Cin2 else None)
    wo, bo = H.t64(w).requires_grad_(True), H.t64(b).requires_grad_(True)
    xin = torch.cat((xo, x2o), 3) if Cin2 else xo
    yo = O.conv2d(xin[:, shift[0]:, shift[1]:, :], wo, bo, stride)
    res = f32(rng.normal(0, 1, tuple(yo.shape))) if residual else None
    reso = H.t64(res).requires_grad_(True) if residual else None
    if residual:
        yo = yo + reso
    dy = f32(rng.normal(0, 1, tuple(yo.shape)))
    yo.backward(H.t64(dy))

    xt, x2t = dev(rt, x), (dev(rt, x2) if Cin2 else None)
    rest = dev(rt, res) if residual else None
    with rt.gradient_tape() as tape:
        y = conv(xt, x2=x2t, residual=rest, shift=shift)
    assert y.shape == tuple(yo.shape)
    seed_grad(rt, y, dy)
    rt.backward(tape)
    check("y", npy(y.data), yo.detach().numpy())
    check("dx", npy(xt.grad), xo.grad.numpy())
    if Cin2:
        check("dx2", npy(x2t.grad), x2o.grad.numpy())
    if residual:
        check("dres", npy(rest.grad), reso.grad.numpy())
    check("dw", npy(conv.kernel.grad), wo.grad.numpy())
    check("db", npy(conv.bias.grad), bo.grad.numpy())


TC_CASES = [
    # (N, H, W, Cin, Cin2, Cout, k, residual): stride-1 shapes the tcgen05 path takes
    (8, 4, 4, 64, 0, 64, 3, False),       # encoder cell conv at the 4x4 scale
    (4, 8, 8, 128, 0, 128, 3, True),      # ... 8x8 scale, with residual (encoder.py:16 style)
    (16, 4, 4, 32, 20, 256, 1, False),    # DecoderSampleCombiner concat(h, z0): K = 52, ragged second source
    (6, 4, 4, 256, 0, 1536, 1, False),    # decoder cell expand 1x1
    (6, 4, 4, 1536, 0, 256, 1, False),    # decoder cell project 1x1
    (6, 8, 8, 128, 0, 40, 3, False),      # sampler head, N = 40
    (2, 16, 16, 64, 0, 96, 5, False),     # postprocess 5x5
    (3, 14, 14, 64, 0, 64, 3, False),     # BASELINE configs[1] micro-bench shapes
    (5, 7, 7, 128, 0, 128, 3, False),
    (2, 32, 32, 32, 0, 32, 3, False),     # W*th tile with two image rows per box
]

S2_CASES = [
    # (N, H, W, Cin, Cout, k, shift, tensor-core directions [fwd, dgrad, wgrad])
    (3, 8, 8, 32, 64, 3, (0, 0), [1, 1, 1]),      # Rescaler DOWN / BNSwishConv stride 2 (common.py:155-163)
    (2, 16, 16, 64, 128, 3, (0, 0), [1, 1, 1]),
    (5, 8, 8, 128, 256, 3, (0, 0), [1, 1, 1]),    # encoder Rescaler at the model's shape
    (2, 8, 8, 32, 16, 1, (0, 0), [1, 0, 1]),      # SkipScaler 1x1 stride 2 (preprocess.py:46-63); dgrad: 3 of 4
    (2, 8, 8, 32, 16, 1, (1, 1), [1, 0, 1]),      # parity classes have no tap -> fp32 CUDA-core path
    (2, 8, 8, 32, 16, 1, (0, 1), [1, 0, 1]),
    (2, 8, 8, 64, 32, 1, (1, 0), [1, 0, 1]),
]


@pytest.mark.parametrize("case", S2_CASES)
def test_conv2d_stride2_tensor_core(lib_built, case):
    """Stride-2 convolutions on tcgen05: x is read through a 5-D TMA view of its four parity planes; backward-data
    runs as four launches, one per output-pixel parity class."""
    import ctypes as C
    from nvae_tf_b200 import _lib
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import Conv2D
    N, Hh, W, Cin, Cout, k, shift, want_tc = case
    rng = np.random.default_rng(6)
    with R.Runtime(seed=7, precision=_lib.NVAE_PREC_TF32X3) as rt:
        conv = Conv2D(Cout, (k, k), strides=(2, 2), padding="same", in_channels=Cin, name="c")
        rt.finalize()
        w = f32(rng.normal(0, 1.0 / np.sqrt(k * k * Cin), (k, k, Cin, Cout)))
        b = f32(rng.normal(0, 0.3, Cout))
        conv.kernel.assign(w)
        conv.bias.assign(b)
        x = f32(rng.normal(0, 1, (N, Hh, W, Cin)))
        xo = H.t64(x).requires_grad_(True)
        wo, bo = H.t64(w).requires_grad_(True), H.t64(b).requires_grad_(True)
        yo = O.conv2d(xo[:, shift[0]:, shift[1]:, :], wo, bo, 2)
        dy = f32(rng.normal(0, 1, tuple(yo.shape)))
        yo.backward(H.t64(dy))
        xt = dev(rt, x)
        d = R.conv_desc(rt, xt.shape, 0, conv.kernel.shape, 2, shift)
        got_tc = [rt.lib._nvae_conv2d_uses_tensor_cores(C.byref(d), i) for i in range(3)]
        assert got_tc == want_tc, got_tc
        with rt.gradient_tape() as tape:
            y = conv(xt, shift=shift)
        assert y.shape == tuple(yo.shape)
        seed_grad(rt, y, dy)
        rt.backward(tape)
        check("y", npy(y.data), yo.detach().numpy())
        check("dx", npy(xt.grad), xo.grad.numpy())
        check("dw", npy(conv.kernel.grad), wo.grad.numpy())
        check("db", npy(conv.bias.grad), bo.grad.numpy())



@pytest.mark.parametrize("mode,tol", [("tf32x3", 2e-5), ("f16x3", 2e-5), ("f16x3-plain", 2e-5), ("tf32", 3e-3)])
@pytest.mark.parametrize("case", TC_CASES + [(8, 16, 16, 384, 0, 384, 5, True)])
def test_conv2d_tensor_core_modes(lib_built, case, mode, tol, monkeypatch):
    """tcgen05 implicit-GEMM fwd / dgrad / wgrad against the float64 oracle.  3xTF32 and 3xFP16 (the arithmetic the
    large GEMMs of NVAE_PREC_TF32X3 use: forced here for every shape, with operand magnitudes far from 1 so the
    absmax scaling matters) must reach the fp32 tolerance; single-pass TF32 is held to the 10-bit-mantissa level."""
    import ctypes as C
    from nvae_tf_b200 import _lib
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import Conv2D
    N, Hh, W, Cin, Cin2, Cout, k, residual = case
    if len(case) == 8 and Cin == 384 and mode == "tf32":
        pytest.skip("large case: 3xTF32 / 3xFP16 only")
    prec = _lib.NVAE_PREC_TF32 if mode == "tf32" else _lib.NVAE_PREC_TF32X3
    monkeypatch.setenv("NVAE_F16X3", "1" if mode.startswith("f16x3") else "0")
    monkeypatch.setenv("NVAE_F16X3_MIN_GFLOP", "0")
    # f16x3: two M tiles per CTA (N <= 192) / two accumulators per tile (N <= 384); -plain: one tile, one accumulator
    monkeypatch.setenv("NVAE_F16X3_DUAL", "0" if mode == "f16x3-plain" else "1")
    monkeypatch.setenv("NVAE_F16X3_NSUB", "0" if mode == "f16x3-plain" else "1")
    xs, dys = (300.0, 1e-5) if mode.startswith("f16x3") else (1.0, 1.0)
    rng = np.random.default_rng(5)
    with R.Runtime(seed=7, precision=prec) as rt:
        conv = Conv2D(Cout, (k, k), padding="same", in_channels=Cin + Cin2, name="c")
        rt.finalize()
        w = f32(rng.normal(0, 1.0 / np.sqrt(k * k * (Cin + Cin2)), (k, k, Cin + Cin2, Cout)))
        b = f32(rng.normal(0, 0.3, Cout))
        conv.kernel.assign(w)
        conv.bias.assign(b)
        x = f32(rng.normal(0, xs, (N, Hh, W, Cin)))
        x2 = f32(rng.normal(0, 1, (N, Hh, W, Cin2))) if Cin2 else None
        xo, x2o = H.t64(x).requires_grad_(True), (H.t64(x2).requires_grad_(True) if Cin2 else None)
        wo, bo = H.t64(w).requires_grad_(True), H.t64(b).requires_grad_(True)
        yo = O.conv2d(torch.cat((xo, x2o), 3) if Cin2 else xo, wo, bo, 1)
        res = f32(rng.normal(0, xs, tuple(yo.shape))) if residual else None
        if residual:
            yo = yo + H.t64(res)
        dy = f32(rng.normal(0, dys, tuple(yo.shape)))
        yo.backward(H.t64(dy))
        xt, x2t = dev(rt, x), (dev(rt, x2) if Cin2 else None)
        d = R.conv_desc(rt, xt.shape, Cin2, conv.kernel.shape, 1)
        assert all(rt.lib._nvae_conv2d_uses_tensor_cores(C.byref(d), i) == 1 for i in range(3)), "not on tcgen05"
        with rt.gradient_tape() as tape:
            y = conv(xt, x2=x2t, residual=dev(rt, res) if residual else None)
        seed_grad(rt, y, dy)
        rt.backward(tape)
        check("y", npy(y.data), yo.detach().numpy(), tol)
        check("dx", npy(xt.grad), xo.grad.numpy(), tol)
        if Cin2:
            check("dx2", npy(x2t.grad), x2o.grad.numpy(), tol)
        check("dw", npy(conv.kernel.grad), wo.grad.numpy(), tol)
        check("db", npy(conv.bias.grad), bo.grad.numpy(), tol)


BN_CONV_CASES = [  # (N, H, W, Cin, Cout, act, residual, bias): the BN -> [swish] -> 1x1 conv pairs of the cells
    (9, 16, 16, 32, 192, 0, False, True),     # decoder conv1 (BN1, no activation)
    (9, 16, 16, 192, 32, 1, False, True),     # decoder conv2 (BN3 + swish)
    (5, 4, 4, 768, 128, 1, True, True),       # deep K, ragged pixel tiles (80 pixels), residual epilogue
    (3, 32, 32, 192, 32, 1, False, False),    # postprocess conv3 (cbs2 BN + swish), no bias
    (7, 8, 8, 64, 384, 0, False, False),      # postprocess cbs1.conv behind bn0
    (144, 4, 4, 256, 1536, 0, False, True),   # bench-size group-0 decoder cell: split-K plans on both launches
    (144, 4, 4, 1536, 256, 2, False, True),   # ... and ELU, for the third activation code
]


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("case", BN_CONV_CASES)
def test_bn_conv2d_operand_prolog(lib_built, case, training, monkeypatch):
    """conv(act(BN(x))) with the BN-apply + activation inside the convolution's operand path (nvae_conv2d_fwd_bnact /
    _wgrad_bnact): against the float64 oracle, and bit for bit against the unfused pair (bn_fwd writes the activated
    tensor, conv2d reads it) -- both feed the same fp32 values to the same tile arithmetic."""
    import ctypes as C
    from nvae_tf_b200 import _lib
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import BatchNormalization, Conv2D
    N, Hh, W, Cin, Cout, act, residual, use_bias = case
    rng = np.random.default_rng(11)
    w = f32(rng.normal(0, 1.0 / np.sqrt(Cin), (1, 1, Cin, Cout)))
    b = f32(rng.normal(0, 0.3, Cout))
    vals = {"gamma": rng.normal(1, 0.2, Cin), "beta": rng.normal(0, 0.3, Cin), "moving_mean": rng.normal(0, 0.5, Cin),
            "moving_variance": rng.uniform(0.5, 2, Cin)}
    x = f32(rng.normal(0.5, 2.0, (N, Hh, W, Cin)))
    res = f32(rng.normal(0, 1, (N, Hh, W, Cout))) if residual else None
    dy = f32(rng.normal(0, 1, (N, Hh, W, Cout)))
    got = {}
    for fused in ("2", "0"):
        monkeypatch.setenv("NVAE_FUSE_BN_CONV", fused)
        with R.Runtime(seed=7, precision=_lib.NVAE_PREC_TF32X3) as rt:
            bn = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=Cin, name="bn")
            conv = Conv2D(Cout, (1, 1), padding="same", use_bias=use_bias, in_channels=Cin, name="c")
            rt.finalize()
            for k, v in vals.items():
                getattr(bn, k).assign(v)
            conv.kernel.assign(w)
            if use_bias:
                conv.bias.assign(b)
            xt = dev(rt, x)
            d = R.conv_desc(rt, xt.shape, 0, conv.kernel.shape, 1)
            assert rt.lib._nvae_conv2d_bnact_supported(C.byref(d)) == 1, "this shape must take the fused path"
            with rt.gradient_tape() as tape:
                y = conv(xt, residual=dev(rt, res) if residual else None, bn_in=(bn, act, training))
            seed_grad(rt, y, dy)
            rt.backward(tape)
            torch.cuda.synchronize()
            got[fused] = dict(y=y.data.clone(), dx=xt.grad.clone(), dw=conv.kernel.grad.clone(),
                              db=conv.bias.grad.clone() if use_bias else None, dgamma=bn.gamma.grad.clone(),
                              dbeta=bn.beta.grad.clone(), mm=bn.moving_mean.value.clone(),
                              mv=bn.moving_variance.value.clone())
    for k in got["2"]:
        if got["2"][k] is not None:
            assert torch.equal(got["2"][k], got["0"][k]), f"{k}: fused and unfused differ"
    # oracle
    p = {"bn/" + k: H.t64(f32(v)) for k, v in vals.items()}
    p["bn/gamma"].requires_grad_(True)
    p["bn/beta"].requires_grad_(True)
    c = O.Ctx(p, training)
    xo = H.t64(x).requires_grad_(True)
    wo, bo = H.t64(w).requires_grad_(True), H.t64(b).requires_grad_(True)
    yo = O.conv2d(ACTS[act](O.batch_norm(c, "bn", xo)), wo, bo if use_bias else None, 1)
    if residual:
        yo = yo + H.t64(res)
    yo.backward(H.t64(dy))
    g = got["2"]
    tol = 5e-5
    check("y", npy(g["y"]), yo.detach().numpy(), tol)
    check("dx", npy(g["dx"]), xo.grad.numpy(), tol)
    check("dw", npy(g["dw"]), wo.grad.numpy(), tol)
    if use_bias:
        check("db", npy(g["db"]), bo.grad.numpy(), tol)
    check("dgamma", npy(g["dgamma"]), p["bn/gamma"].grad.numpy(), tol)
    check("dbeta", npy(g["dbeta"]), p["bn/beta"].grad.numpy(), tol)
    if training:
        check("moving_mean", npy(g["mm"]), c.new_stats["bn/moving_mean"].numpy())
        check("moving_variance", npy(g["mv"]), c.new_stats["bn/moving_variance"].numpy())


@pytest.mark.parametrize("shape", [(144, 16, 16, 384, 384), (144, 32, 32, 192, 192)])
def test_conv2d_full_size_3xfp16_agrees_with_3xtf32(lib_built, shape, monkeypatch):
    """BASELINE-size check of the dominant GEMMs (batch 144; the float64 oracle would need minutes per case): the 3xFP16
    path the step uses for them (two accumulators / two M tiles per CTA, pixel-aligned split of the 226 MB filter
    gradient) must agree with the independent 3xTF32 path on the same operands -- forward, backward-data and
    backward-filter, each far inside the 1e-3 bound -- and 5x5 SAME convolution is linear: conv(2x, w) == 2 conv(x, w)
    bit for bit (power-of-two scale: the absmax-derived operand scale moves by exactly one exponent)."""
    import ctypes as C
    from nvae_tf_b200 import _lib
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import Conv2D
    N, Hh, W, Cin, Cout = shape
    monkeypatch.delenv("NVAE_F16X3_MIN_GFLOP", raising=False)
    g = torch.Generator(device="cpu").manual_seed(11)
    with R.Runtime(seed=7, precision=_lib.NVAE_PREC_TF32X3) as rt:
        conv = Conv2D(Cout, (5, 5), padding="same", in_channels=Cin, name="c")
        rt.finalize()
        conv.kernel.assign((torch.randn(5, 5, Cin, Cout, generator=g) / np.sqrt(25 * Cin)).numpy())
        conv.bias.assign(torch.zeros(Cout).numpy())
        x = torch.randn(N, Hh, W, Cin, generator=g)
        x = x * torch.rand(1, 1, 1, Cin, generator=g) * 30.0          # per-channel magnitudes spread over a decade+
        dy = torch.randn(N, Hh, W, Cout, generator=g) * 1e-4
        res = {}
        # three arithmetics on the same operands: 3xTF32; 3xFP16 (three-term); the step's default for these >= 20 GFLOP
        # launches, the two-term 3xFP16 product (B rounded to fp16; NVAE_F16X2=0 switches it off)
        MODES = {"tf32x3": ("0", "0"), "f16x3": ("1", "0"), "f16x2": ("1", "1")}
        for mode, (f16, two) in MODES.items():
            monkeypatch.setenv("NVAE_F16X3", f16)
            monkeypatch.setenv("NVAE_F16X2", two)
            d = R.conv_desc(rt, tuple(x.shape), 0, conv.kernel.shape, 1)
            info = (C.c_int32 * 16)()
            rt.lib._nvae_conv2d_plan_info(C.byref(d), 0, info)
            assert info[9] == int(f16)  # the arithmetic under test is the one that runs
            xt = R.DeviceTensor(x.to(rt.device), True)
            with rt.gradient_tape() as tape:
                y = conv(xt)
            y.grad = dy.to(rt.device)
            rt.backward(tape)
            torch.cuda.synchronize()
            res[mode] = (y.data.double().cpu(), xt.grad.double().cpu(), torch.as_tensor(conv.kernel.grad).double().cpu().clone())
            if f16 == "1":
                x2 = R.DeviceTensor((2.0 * x).to(rt.device), False)
                y2 = conv(x2)
                torch.cuda.synchronize()
                assert torch.equal(y2.data.cpu(), 2.0 * y.data.cpu())
        assert not torch.equal(res["f16x3"][0], res["f16x2"][0])  # the switch really selects another arithmetic
        # the two three-term arithmetics against each other (3xTF32 splits by truncation: its dropped terms are biased and add
        # up over K = 9600; 3xFP16 rounds to nearest)
        for name, a, b in zip(("y", "dx", "dw"), res["f16x3"], res["tf32x3"]):
            err = float((a - b).abs().max() / b.abs().max())
            assert err <= 2e-4, f"{name}: 3xFP16 vs 3xTF32 max|d|/max|ref| = {err:.2e}"
        # ... and all three against the float64 oracle on the first image (forward and backward-data are per-sample).  At
        # K = 9600 the floor of the three-term products is the tensor core's own fp32 accumulation over 1 800 - 3 600 chained
        # MMAs (~3e-5 measured); the two-term product adds the fp16 rounding of B (2^-12 per weight, unbiased)
        TOLS = {"f16x3": (1e-4, 1e-4), "tf32x3": (2e-4, 4e-4), "f16x2": (5e-4, 5e-4)}  # (y / dx, dw)
        xo = x[:1].double().requires_grad_(True)
        yo = O.conv2d(xo, torch.as_tensor(npy(conv.kernel.value)), None, 1)
        yo.backward(dy[:1].double())
        worst = {}
        for mode, (tol, _) in TOLS.items():
            ey = float((res[mode][0][:1] - yo.detach()).abs().max() / yo.detach().abs().max())
            ex = float((res[mode][1][:1] - xo.grad).abs().max() / xo.grad.abs().max())
            worst[mode] = [ey, ex]
            assert ey <= tol and ex <= tol, f"{mode}: y {ey:.2e}, dx {ex:.2e} vs float64"
        # the full-size FILTER gradient of all three against float64 (batch 144: the pixel-aligned split-K, nsub /
        # dual tiles).  dw[r,s,ci,co] = sum_pixels x[n,h+r-2,w+s-2,ci] * dy[n,h,w,co]: a float64 GEMM on the device over
        # ALL pixels for a sample of taps (corners, centre, an edge) and 16 input channels each -- 6 144 (or 3 072) of the
        # filter's entries per tap, every one a sum over all 36 864 / 147 456 pixels
        xd, dyd = x.to(rt.device).double(), dy.to(rt.device).double()
        xp = torch.nn.functional.pad(xd, (0, 0, 2, 2, 2, 2))  # SAME padding of a 5x5 stride-1 conv: 2 before, 2 after
        ci = torch.arange(0, Cin, Cin // 16, device=rt.device)[:16]
        dyf = dyd.reshape(-1, Cout)
        edw = {m: 0.0 for m in TOLS}
        for (r, sx) in ((0, 0), (2, 2), (4, 4), (1, 3), (4, 0)):
            xs = xp[:, r:r + Hh, sx:sx + W, :][..., ci].reshape(-1, ci.numel())
            ref = (xs.t() @ dyf).cpu()                                    # [16, Cout] float64
            for mode, (_, tol) in TOLS.items():
                got = res[mode][2][r, sx][ci.cpu()]
                err = float((got - ref).abs().max() / ref.abs().max())
                edw[mode] = max(edw[mode], err)
                assert err <= tol, f"{mode}: dw tap ({r},{sx}) vs float64: {err:.2e}"
        # forward over ALL images against a float64 device GEMM for a sample of output channels (im2col of 16 columns)
        wv = torch.as_tensor(npy(conv.kernel.value)).to(rt.device)       # [5,5,Cin,Cout] float64
        co = torch.arange(0, Cout, Cout // 8, device=rt.device)[:8]
        yref = torch.zeros(N, Hh, W, co.numel(), dtype=torch.float64, device=rt.device)
        for r in range(5):
            for sx in range(5):
                yref += xp[:, r:r + Hh, sx:sx + W, :] @ wv[r, sx][:, co]
        for mode, (tol, _) in TOLS.items():
            got = res[mode][0][..., co.cpu()]
            err = float((got - yref.cpu()).abs().max() / yref.abs().max())
            assert err <= tol, f"{mode}: y (all images) vs float64: {err:.2e}"
        print(f"full-size {shape}: max rel err vs float64 [y, dx] / dw: " +
              "; ".join(f"{m} {worst[m][0]:.1e} {worst[m][1]:.1e} / {edw[m]:.1e}" for m in TOLS))


def test_conv2d_concat_output_and_accumulating_dgrad(rt):
    """SkipScaler (preprocess.py:65-74): four strided 1x1 convs on shifted views write channel slices of one
    tensor; their dgrads accumulate into the shared input gradient."""
    from nvae_tf_b200.preprocess import SkipScaler
    rng = np.random.default_rng(2)
    sk = SkipScaler(16, in_channels=8, name="skip")
    rt.finalize()
    p = {}
    for v in rt.variables.values():
        a = f32(rng.normal(0, 0.3, v.shape))
        v.assign(a)
        p[v.name] = H.t64(a)
    x = f32(rng.normal(0, 1, (2, 8, 8, 8)))
    xt = dev(rt, x)
    with rt.gradient_tape() as tape:
        y = sk(xt, training=False)
    xo = H.t64(x).requires_grad_(True)
    c = O.Ctx(p, False)
    o = O.swish(xo)
    yo = torch.cat((O.sn_conv(c, "skip/conv1", o, 2), O.sn_conv(c, "skip/conv2", o[:, 1:, 1:, :], 2),
                    O.sn_conv(c, "skip/conv3", o[:, :, 1:, :], 2), O.sn_conv(c, "skip/conv4", o[:, 1:, :, :], 2)), 3)
    dy = f32(rng.normal(0, 1, tuple(yo.shape)))
    yo.backward(H.t64(dy))
    seed_grad(rt, y, dy)
    rt.backward(tape)
    check("y", npy(y.data), yo.detach().numpy())
    check("dx", npy(xt.grad), xo.grad.numpy())


@pytest.mark.parametrize("shape", [(3, 4, 4, 192), (2, 8, 8, 96), (2, 7, 5, 32), (1, 14, 14, 64)])
@pytest.mark.parametrize("training", [True, False])
def test_dwconv_bn_swish(rt, shape, training):
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import BatchNormalization, DepthwiseConv2D
    rng = np.random.default_rng(3)
    C = shape[-1]
    bn = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=C, name="bn")
    dw = DepthwiseConv2D((5, 5), padding="same", in_channels=C, name="dw")
    rt.finalize()
    p = {}
    for v in rt.variables.values():
        a = f32(rng.uniform(0.5, 1.5, v.shape) if "variance" in v.name else rng.normal(0, 0.4, v.shape))
        v.assign(a)
        p[v.name] = H.t64(a).requires_grad_(v.trainable)
    x = f32(rng.normal(0, 1.5, shape))
    dy = f32(rng.normal(0, 1, shape))
    xt = dev(rt, x)
    with rt.gradient_tape() as tape:
        y = R.dwconv_bn_act(rt, xt, bn, 1, dw, training)
    seed_grad(rt, y, dy)
    rt.backward(tape)
    xo = H.t64(x).requires_grad_(True)
    c = O.Ctx(p, training)
    yo = O.depthwise_conv2d(O.swish(O.batch_norm(c, "bn", xo)), p["dw/depthwise_kernel"], p["dw/bias"])
    yo.backward(H.t64(dy))
    check("y", npy(y.data), yo.detach().numpy())
    check("dx", npy(xt.grad), xo.grad.numpy())
    check("dw", npy(dw.depthwise_kernel.grad), p["dw/depthwise_kernel"].grad.numpy())
    check("dbias", npy(dw.bias.grad), p["dw/bias"].grad.numpy())
    check("dgamma", npy(bn.gamma.grad), p["bn/gamma"].grad.numpy())
    check("dbeta", npy(bn.beta.grad), p["bn/beta"].grad.numpy())


@pytest.mark.parametrize("shape,with_bn,alpha,beta", [((5, 4, 4, 64), False, 0.1, 1.0), ((3, 8, 8, 32), True, 0.1, 1.0),
                                                      ((2, 16, 16, 8), True, 1.0, 0.1), ((4, 7, 7, 128), False, 1.0, 0.1),
                                                      ((2, 2, 2, 512), True, 0.1, 1.0)])
def test_se_residual(rt, shape, with_bn, alpha, beta):
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.common import SqueezeExcitation
    from nvae_tf_b200.layers import BatchNormalization
    rng = np.random.default_rng(4)
    C = shape[-1]
    se = SqueezeExcitation(channels=C, name="se")
    bn = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=C, name="bn") if with_bn else None
    rt.finalize()
    assert se.dense1.units == int(max(C / 16, 4))
    p = {}
    for v in rt.variables.values():
        a = f32(rng.uniform(0.5, 1.5, v.shape) if "variance" in v.name else rng.normal(0, 0.5, v.shape))
        v.assign(a)
        p[v.name] = H.t64(a).requires_grad_(v.trainable)
    t, xr, dy = (f32(rng.normal(0, 1, shape)) for _ in range(3))
    tt, xrt = dev(rt, t), dev(rt, xr)
    with rt.gradient_tape() as tape:
        y = R.se_residual(rt, tt, bn, xrt, se, alpha, beta, True)
    seed_grad(rt, y, dy)
    rt.backward(tape)
    to, xro = H.t64(t).requires_grad_(True), H.t64(xr).requires_grad_(True)
    c = O.Ctx(p, True)
    u = O.batch_norm(c, "bn", to) if with_bn else to
    yo = alpha * xro + beta * O.squeeze_excitation(c, "se", u)
    yo.backward(H.t64(dy))
    check("y", npy(y.data), yo.detach().numpy())
    check("dt", npy(tt.grad), to.grad.numpy())
    check("dxres", npy(xrt.grad), xro.grad.numpy())
    for n in ("dense1/kernel", "dense1/bias", "dense2/kernel", "dense2/bias"):
        check(n, npy(rt.variables["se/" + n].grad), p["se/" + n].grad.numpy())
    if with_bn:
        check("dgamma", npy(bn.gamma.grad), p["bn/gamma"].grad.numpy())
        check("dbeta", npy(bn.beta.grad), p["bn/beta"].grad.numpy())


@pytest.mark.parametrize("B,HW,L,has_dec", [(5, 16, 20, True), (4, 64, 20, True), (3, 16, 20, False), (2, 1, 3, True),
                                             (144, 16, 20, True)])
def test_latent_fwd_bwd(rt, B, HW, L, has_dec):
    rng = np.random.default_rng(5)
    rt.finalize()
    h = int(round(HW ** 0.5))
    shp, shp2 = (B, h, HW // h, L), (B, h, HW // h, 2 * L)
    enc, dec, eps, dz = f32(rng.normal(0, 2, shp2)), f32(rng.normal(0, 2, shp2)), f32(rng.normal(0, 1, shp)), \
        f32(rng.normal(0, 1, shp))
    w = 0.37
    lib = rt.lib
    te, td, tz = (torch.as_tensor(a.astype(np.float32)).to(rt.device) for a in (enc, dec, eps))
    z, dist = rt.empty(*shp), rt.empty(4, *shp)
    kl, lq, lp = rt.empty(B), rt.zeros(B), rt.zeros(B)
    lib.latent_fwd(te.data_ptr(), td.data_ptr() if has_dec else None, tz.data_ptr(), B, HW, L, z.data_ptr(),
                   kl.data_ptr(), lq.data_ptr(), lp.data_ptr(), dist.data_ptr(), rt.stream)
    klw = torch.tensor([w], device=rt.device)
    tdz = torch.as_tensor(dz.astype(np.float32)).to(rt.device)
    de, dd = rt.empty(*shp2), rt.empty(*shp2)
    lib.latent_bwd(te.data_ptr(), td.data_ptr() if has_dec else None, tz.data_ptr(), tdz.data_ptr(), klw.data_ptr(), B,
                   HW, L, de.data_ptr(), dd.data_ptr() if has_dec else None, rt.stream)
    # oracle: common.py:76-102 + models.py:197-201 + util.py:39-46
    eo, do = H.t64(enc).requires_grad_(True), H.t64(dec).requires_grad_(True)
    a, b = torch.chunk(eo, 2, -1)
    if has_dec:
        cm, cs = torch.chunk(do, 2, -1)
        mu_p, sig_p = O.softclamp5(cm), torch.exp(O.softclamp5(cs)) + 1e-2
        mu_q, sig_q = O.softclamp5(a + cm), torch.exp(O.softclamp5(cs + b)) + 1e-2
    else:
        mu_q, sig_q = O.softclamp5(a), torch.exp(O.softclamp5(b)) + 1e-2
        mu_p, sig_p = torch.zeros_like(mu_q), torch.ones_like(sig_q)
    zo = mu_q + H.t64(eps) * sig_q
    klo = O.kl_per_group([O.DistributionParams(mu_q, sig_q, mu_p, sig_p)])[0]
    (w * klo.sum() + (zo * H.t64(dz)).sum()).backward()
    check("z", npy(z), zo.detach().numpy())
    check("kl", npy(kl), klo.detach().numpy())
    check("log_q", npy(lq), O.calculate_log_p(zo, mu_q, sig_q).sum(dim=(1, 2, 3)).detach().numpy(), tol=5e-5)
    check("log_p", npy(lp), O.calculate_log_p(zo, mu_p, sig_p).sum(dim=(1, 2, 3)).detach().numpy(), tol=5e-5)
    for i, ref in enumerate((mu_q, sig_q, mu_p, sig_p)):
        check(f"dist{i}", npy(dist[i]), ref.detach().numpy())
    check("d_enc", npy(de), eo.grad.numpy(), tol=5e-5)
    if has_dec:
        check("d_dec", npy(dd), do.grad.numpy(), tol=5e-5)


@pytest.mark.parametrize("B,C,Cl,crop", [(6, 1, 1, 0), (3, 1, 1, 2), (2, 3, 1, 0), (2, 3, 3, 0)])
def test_bernoulli_ll(rt, B, C, Cl, crop):
    rng = np.random.default_rng(6)
    rt.finalize()
    l = f32(rng.normal(0, 4, (B, 32, 32, Cl)))
    x = (rng.random((B, 32, 32, C)) < 0.3).astype(np.float64)
    tl, tx = (torch.as_tensor(a.astype(np.float32)).to(rt.device) for a in (l, x))
    out, dl = rt.empty(B), rt.empty(B, 32, 32, Cl)
    rt.lib.bernoulli_ll_fwd(tl.data_ptr(), tx.data_ptr(), B, 32, 32, C, Cl, crop, out.data_ptr(), rt.stream)
    lo = H.t64(l).requires_grad_(True)
    ro = O.calculate_recon_loss(H.t64(x), lo, crop_output=bool(crop))
    check("recon", npy(out), ro.detach().numpy())
    if not crop:
        rt.lib.bernoulli_ll_bwd(tl.data_ptr(), tx.data_ptr(), B, 32, 32, C, Cl, 1.0 / B, dl.data_ptr(), rt.stream)
        ro.mean().backward()
        check("dlogits", npy(dl), lo.grad.numpy())


def test_loss_assemble_and_bn_loss(rt):
    from nvae_tf_b200.layers import BatchNormalization
    rng = np.random.default_rng(7)
    bns = [BatchNormalization(momentum=0.05, epsilon=1e-5, channels=c, name=f"bn{i}", in_bn_loss=True)
           for i, c in enumerate([32, 64, 128, 1536, 4])]
    rt.finalize()
    gam = []
    for bn in bns:
        g = f32(rng.normal(1, 0.5, bn.gamma.shape))
        g[1] = g[0] = -np.abs(g).max() - 0.25  # a tie between two negative maxima (tf.reduce_max gradient splits it)
        bn.gamma.assign(g)
        gam.append(H.t64(g).requires_grad_(True))
    loss = rt.zeros(1)
    rt.lib.bn_loss_fwd(rt.params.data_ptr(), rt.bn_loss_offsets.data_ptr(), rt.bn_loss_sizes.data_ptr(), rt.bn_loss_n,
                       0.01, loss.data_ptr(), rt.stream)
    rt.lib.fill(rt.grads.data_ptr(), rt.grads.numel(), 0.0, rt.stream)
    rt.lib.bn_loss_bwd(rt.params.data_ptr(), rt.grads.data_ptr(), rt.bn_loss_offsets.data_ptr(),
                       rt.bn_loss_sizes.data_ptr(), rt.bn_loss_n, 0.01, rt.stream)
    ref = 0.01 * sum(g.abs().amax() for g in gam)  # amax: evenly split sub-gradient, as tf.reduce_max
    ref.backward()
    check("bn_loss", npy(loss), ref.detach().numpy().reshape(1))
    for bn, g in zip(bns, gam):
        check("dgamma", npy(bn.gamma.grad), g.grad.numpy())
    # KL balancing + loss assembly (models.py:121-126, 204-222)
    cfg = O.NVAEConfig()
    G, B = 15, 9
    kl_all, recon = f32(rng.uniform(0, 30, (G, B))), f32(rng.uniform(50, 100, B))
    alphas = O.kl_alphas(cfg)
    for beta in (0.0, 0.4, 1.0):
        hyper = torch.tensor([beta] + [0] * 7, dtype=torch.float32, device=rt.device)
        tk, tr, ta = (torch.as_tensor(np.asarray(a, np.float32)).to(rt.device) for a in (kl_all, recon, alphas))
        klw, kll, sc = rt.empty(G), rt.empty(B), rt.empty(2)
        rt.lib.loss_assemble(tk.data_ptr(), tr.data_ptr(), loss.data_ptr(), ta.data_ptr(), hyper.data_ptr(), -1, G, B,
                             klw.data_ptr(), kll.data_ptr(), sc.data_ptr(), rt.stream)
        ko = H.t64(kl_all).requires_grad_(True)
        zp = None
        if beta < 1:
            coeff = ko.abs().mean(1) + 0.01
            coeff = coeff / H.t64(alphas) * coeff.sum()
            coeff = (coeff / coeff.mean()).detach()
            kl = (ko * coeff[:, None]).sum(0)
        else:
            kl = ko.sum(0)
        tot = (H.t64(recon) + beta * kl).mean() + ref.detach()
        tot.backward()
        check("kl_loss", npy(kll), (beta * kl).detach().numpy(), floor=1e-3)
        check("total", npy(sc[:1]), tot.detach().numpy().reshape(1))
        check("kl_weight", npy(klw), ko.grad.numpy()[:, 0], floor=1e-6)


def test_spectral_norm_all_layers_one_pass(rt):
    from nvae_tf_b200.layers import Conv2D, SpectralNormalization
    rng = np.random.default_rng(8)
    shapes = [(3, 3, 1, 32), (1, 1, 276, 256), (5, 5, 48, 48), (3, 3, 128, 40), (1, 1, 64, 16), (3, 3, 256, 256)]
    sns = [SpectralNormalization(Conv2D(s[3], (s[0], s[1]), padding="same", in_channels=s[2], name=f"c{i}"))
           for i, s in enumerate(shapes)]
    rt.finalize()
    ws, us = [], []
    for sn in sns:
        w, u = f32(rng.normal(0, 0.2, sn.layer.kernel.shape)), f32(rng.normal(0, 0.02, sn.u.shape))
        sn.layer.kernel.assign(w)
        sn.u.assign(u)
        ws.append(w)
        us.append(u)
    rt.spectral_normalize_all()
    for i, sn in enumerate(sns):
        c = O.Ctx({"c/kernel": H.t64(ws[i]), "c/u": H.t64(us[i])}, True)
        wn = O.sn_kernel(c, "c")
        check(f"w{i}", npy(sn.layer.kernel.value), wn.numpy())
        check(f"u{i}", npy(sn.u.value), c.new_stats["c/u"].numpy())
    # a second iteration through the per-layer entry (layers called outside NVAE.call)
    w1, u1 = npy(sns[2].layer.kernel.value), npy(sns[2].u.value)
    rt.spectral_normalize_one(2)
    c = O.Ctx({"c/kernel": H.t64(w1), "c/u": H.t64(u1)}, True)
    check("w second", npy(sns[2].layer.kernel.value), O.sn_kernel(c, "c").numpy())
    check("others untouched", npy(sns[3].layer.kernel.value), npy(sns[3].layer.kernel.value))


def test_adamax_cosine_schedule_and_beta(rt):
    rng = np.random.default_rng(9)
    rt.finalize()
    n = 4096
    p0, g0 = f32(rng.normal(0, 1, n)), f32(rng.normal(0, 1, n))
    p = torch.as_tensor(p0.astype(np.float32)).to(rt.device)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    counters = torch.tensor([7, 0], dtype=torch.int64, device=rt.device)
    hyper = torch.zeros(8, device=rt.device)
    po, mo, vo = H.t64(p0), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    for it in range(3):
        g = f32(g0 * (it + 1))
        tg = torch.as_tensor(g.astype(np.float32)).to(rt.device)
        rt.lib.schedule_step(counters.data_ptr(), hyper.data_ptr(), 0.3 * 100, 1e-3, 50.0, 0.9, 3, rt.stream)
        rt.lib.adamax(p.data_ptr(), tg.data_ptr(), m.data_ptr(), v.data_ptr(), n, hyper.data_ptr(), 0.9, 0.999, 1e-7,
                      0.5, rt.stream)
        lr = O.cosine_decay_lr(it, 50)
        po, mo, vo = O.adamax_update(po, 0.5 * H.t64(g), mo, vo, it + 1, lr)
        assert abs(hyper[0].item() - min((7 + it) / 30.0, 1.0)) < 1e-6
    assert counters.tolist() == [10, 3]
    check("p", npy(p), po.numpy(), tol=1e-6)
    check("m", npy(m), mo.numpy(), tol=1e-6)
    check("v", npy(v), vo.numpy(), tol=1e-6)


def test_philox_normal_moments_and_streams(rt):
    rt.finalize()
    a, b = rt.empty(1 << 20), rt.empty(1 << 20)
    rt.lib.philox_normal(a.data_ptr(), a.numel(), 1234, None, 0, rt.stream)
    rt.lib.philox_normal(b.data_ptr(), b.numel(), 1234, None, 1, rt.stream)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1) < 5e-3
    assert abs((a * b).mean().item()) < 5e-3 and not torch.equal(a, b)
    c = rt.empty(1 << 20)
    rt.lib.philox_normal(c.data_ptr(), c.numel(), 1234, None, 0, rt.stream)
    assert torch.equal(a, c)  # counter-based: reproducible


@pytest.mark.parametrize("shape", [(144, 16, 16, 384, 384), (144, 32, 32, 192, 192), (144, 16, 16, 64, 384)])
def test_large_convolutions_are_bitwise_repeatable(lib_built, shape, monkeypatch):
    """Forward, backward-data and backward-filter of the step's large GEMMs, 30 launches back to back with the workspace
    scrambled in between, must give bit-identical results: split-K partials are summed in a fixed order and nothing may
    depend on what the workspace held or on kernel timing.  (Round 2: this caught the N = 384 backward-filter reading a
    tile that had not landed -- TMEM A ring one slot deeper than the shared-memory ring -- in ~1 launch of 13.)"""
    import ctypes as C
    from nvae_tf_b200 import _lib
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.layers import Conv2D
    monkeypatch.delenv("NVAE_F16X3_MIN_GFLOP", raising=False)
    N, Hh, W, Cin, Cout = shape
    k = 5 if Cin == Cout else 1
    g = torch.Generator(device="cpu").manual_seed(5)
    with R.Runtime(seed=7, precision=_lib.NVAE_PREC_TF32X3) as rt:
        conv = Conv2D(Cout, (k, k), padding="same", in_channels=Cin, name="c")
        rt.finalize()
        rt.pack_plain(conv)
        x = torch.randn(N, Hh, W, Cin, generator=g).to(rt.device)
        dy = (torch.randn(N, Hh, W, Cout, generator=g) * 1e-3).to(rt.device)
        d = R.conv_desc(rt, tuple(x.shape), 0, conv.kernel.shape, 1)
        ws, wsb = rt.workspace(max(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), i) for i in range(3)))
        y, dx = torch.empty(N, Hh, W, Cout, device=rt.device), torch.empty(N, Hh, W, Cin, device=rt.device)
        ref = None
        for rep in range(30):
            rt._ws["main"].random_(0, 255)
            rt.lib.conv2d_fwd(C.byref(d), x.data_ptr(), None, conv.kernel.ptr(), conv.packed_fwd(), None, None, y.data_ptr(),
                              ws, wsb, rt.stream)
            rt.lib.conv2d_dgrad(C.byref(d), dy.data_ptr(), conv.kernel.ptr(), conv.packed_dgrad(), dx.data_ptr(), None, 0, ws,
                                wsb, rt.stream)
            rt.lib.conv2d_wgrad(C.byref(d), x.data_ptr(), None, dy.data_ptr(), conv.kernel.gptr(), None, ws, wsb, rt.stream)
            out = (y.clone(), dx.clone(), conv.kernel.grad.clone())
            if ref is None:
                ref = out
            else:
                for name, a, b in zip(("y", "dx", "dw"), ref, out):
                    assert torch.equal(a, b), f"{name} differs on launch {rep}"
        torch.cuda.synchronize()
