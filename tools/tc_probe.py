#!/usr/bin/env python
"""GPU probe of the tcgen05 convolution path (development tool, not a test):
  1. does kind::tf32 truncate or round raw fp32 operand bits?
  2. forward / dgrad / wgrad of R.conv2d in TF32 mode against a float64 reference, on TF32-exact inputs
     (isolates layout bugs: the only error left is fp32 accumulation order) and on raw fp32 inputs
     (the real TF32 error level).
Run on the GPU box:  python tools/tc_probe.py [quick]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvae_tf_b200 import _lib  # noqa: E402
from nvae_tf_b200 import runtime as R  # noqa: E402
from nvae_tf_b200.layers import Conv2D  # noqa: E402


PREC = {"tf32": _lib.NVAE_PREC_TF32, "tf32x3": _lib.NVAE_PREC_TF32X3}[os.environ.get("NVAE_PRECISION", "tf32x3")]


def tf32_round(a: np.ndarray) -> np.ndarray:
    """round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32.f32"""
    u = np.asarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def ref_conv(x, w, b, dy):
    """float64 conv (stride 1, SAME, odd kernels) + grads on the GPU via torch (NCHW internally)."""
    xt = torch.as_tensor(x, dtype=torch.float64, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    wt = torch.as_tensor(w, dtype=torch.float64, device="cuda").permute(3, 2, 0, 1).requires_grad_(True)
    bt = torch.as_tensor(b, dtype=torch.float64, device="cuda").requires_grad_(True)
    y = torch.nn.functional.conv2d(xt, wt, bt, padding=(w.shape[0] // 2, w.shape[1] // 2))
    y.backward(torch.as_tensor(dy, dtype=torch.float64, device="cuda").permute(0, 3, 1, 2))
    return (y.detach().permute(0, 2, 3, 1).cpu().numpy(), xt.grad.permute(0, 2, 3, 1).cpu().numpy(),
            wt.grad.permute(2, 3, 1, 0).cpu().numpy(), bt.grad.cpu().numpy())


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run_case(N, H, W, Cin, Cin2, Cout, k, exact, seed=0, time_it=False):
    rng = np.random.default_rng(seed)
    rt = R.Runtime(precision=PREC, seed=1)
    with rt:
        conv = Conv2D(Cout, (k, k), padding="same", in_channels=Cin + Cin2, name="c")
        rt.finalize()
        w = rng.normal(0, 1.0 / np.sqrt(k * k * (Cin + Cin2)), (k, k, Cin + Cin2, Cout)).astype(np.float32)
        b = rng.normal(0, 0.3, Cout).astype(np.float32)
        x = rng.normal(0, 1, (N, H, W, Cin + Cin2)).astype(np.float32)
        dy = rng.normal(0, 1, (N, H, W, Cout)).astype(np.float32)
        if exact:
            w, x, dy = tf32_round(w), tf32_round(x), tf32_round(dy)
        conv.kernel.assign(w)
        conv.bias.assign(b)
        xt = R.DeviceTensor(torch.as_tensor(np.ascontiguousarray(x[..., :Cin])).cuda())
        x2t = R.DeviceTensor(torch.as_tensor(np.ascontiguousarray(x[..., Cin:])).cuda()) if Cin2 else None
        d = R.conv_desc(rt, xt.shape, Cin2, conv.kernel.shape, 1)
        import ctypes as C
        tc = [rt.lib._nvae_conv2d_uses_tensor_cores(C.byref(d), i) for i in range(3)]
        with rt.gradient_tape() as tape:
            y = conv(xt, x2=x2t)
        y.grad = torch.as_tensor(dy).cuda()
        rt.backward(tape)
        torch.cuda.synchronize()
        yo, dxo, dwo, dbo = ref_conv(x, w, b, dy)
        dx = xt.grad.cpu().numpy()
        if Cin2:
            dx = np.concatenate((dx, x2t.grad.cpu().numpy()), 3)
        errs = (rel(y.data.cpu().numpy(), yo), rel(dx, dxo), rel(conv.kernel.grad.cpu().numpy(), dwo),
                rel(conv.bias.grad.cpu().numpy(), dbo))
        line = (f"N{N} {H}x{W} Cin{Cin}+{Cin2} Cout{Cout} k{k} exact={int(exact)} tc={tc}: "
                f"y {errs[0]:.2e} dx {errs[1]:.2e} dw {errs[2]:.2e} db {errs[3]:.2e}")
        if time_it:
            ws, wsb = rt.workspace(max(rt.lib._nvae_conv2d_ws_bytes(C.byref(d), i) for i in range(3)))
            dyt = torch.as_tensor(dy).cuda()
            dxb = torch.empty_like(xt.data)
            fl = 2.0 * N * H * W * Cout * k * k * (Cin + Cin2)
            for name, fn in (
                ("fwd", lambda: rt.lib.conv2d_fwd(C.byref(d), xt.ptr(), None, conv.kernel.ptr(), conv.packed_fwd(),
                                                  conv.bias.ptr(), None, y.ptr(), ws, wsb, rt.stream)),
                ("dgrad", lambda: rt.lib.conv2d_dgrad(C.byref(d), dyt.data_ptr(), conv.kernel.ptr(),
                                                      conv.packed_dgrad(), dxb.data_ptr(), None, 0, ws, wsb,
                                                      rt.stream)),
                ("wgrad", lambda: rt.lib.conv2d_wgrad(C.byref(d), xt.ptr(), None, dyt.data_ptr(), conv.kernel.gptr(),
                                                      None, ws, wsb, rt.stream))):
                if Cin2:
                    break
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                line += f" | {name} {ms * 1e3:.0f}us {fl / ms / 1e9:.0f}TF/s"
        print(line, flush=True)
        return errs


def rounding_probe():
    rt = R.Runtime(precision=_lib.NVAE_PREC_TF32, seed=1)
    with rt:
        conv = Conv2D(32, (1, 1), padding="same", in_channels=32, use_bias=False, name="c")
        rt.finalize()
        w = np.zeros((1, 1, 32, 32), np.float32)
        w[0, 0, 0, :] = 1.0
        conv.kernel.assign(w)
        x = np.zeros((8, 4, 4, 32), np.float32)
        vals = [1 + 3 * 2.0 ** -12, 1 + 1 * 2.0 ** -12, 1 + 2.0 ** -11, 1 + 2.0 ** -10 + 2.0 ** -11, -(1 + 3 * 2.0 ** -12)]
        for i, v in enumerate(vals):
            x[i, :, :, 0] = v
        rt.tf32_round = "none"
        xt = R.DeviceTensor(torch.as_tensor(x).cuda(), needs_grad=False)
        y = conv(xt)
        torch.cuda.synchronize()
        out = y.data.cpu().numpy()[:len(vals), 0, 0, 0]
        for v, o in zip(vals, out):
            kind = "exact-fp32" if o == np.float32(v) else "RN" if o == tf32_round(np.float32(v)) else \
                "truncate" if o == np.float32(np.sign(v) * np.floor(abs(v) * 1024) / 1024) else "?"
            print(f"rounding probe: in {v!r} -> out {float(o)!r}  [{kind}]", flush=True)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    t0 = time.time()
    rounding_probe()
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = [(8, 4, 4, 32, 0, 32, 1), (8, 4, 4, 64, 0, 64, 3), (4, 8, 8, 128, 0, 128, 3), (2, 16, 16, 64, 0, 96, 5),
             (2, 32, 32, 32, 0, 32, 3), (16, 4, 4, 32, 20, 256, 1), (16, 4, 4, 256, 20, 256, 1), (6, 8, 8, 128, 0, 40, 3),
             (6, 4, 4, 256, 0, 1536, 1), (6, 4, 4, 1536, 0, 256, 1), (3, 14, 14, 64, 0, 64, 3), (5, 7, 7, 128, 0, 128, 3),
             (2, 16, 16, 384, 0, 384, 5)]
    for c in cases:
        run_case(*c, exact=True)
    for c in cases[:6]:
        run_case(*c, exact=False)
    if not quick:
        for c in [(144, 16, 16, 384, 0, 384, 5), (144, 32, 32, 192, 0, 192, 5), (144, 8, 8, 128, 0, 128, 3),
                  (144, 4, 4, 256, 0, 256, 3), (144, 4, 4, 256, 0, 1536, 1), (144, 8, 8, 768, 0, 128, 1),
                  (144, 32, 32, 32, 0, 192, 1), (144, 32, 32, 192, 0, 32, 1)]:
            run_case(*c, exact=True, time_it=True)
    print(f"probe done in {time.time() - t0:.1f}s")
