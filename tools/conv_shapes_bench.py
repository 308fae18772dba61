#!/usr/bin/env python
"""Every distinct convolution shape of the default NVAE step, timed per direction (fwd / dgrad / wgrad incl. its
fix-up launch) from a CUDA graph of back-to-back launches, with its multiplicity in the step.
usage (GPU box): python tools/conv_shapes_bench.py [batch] > gpurun_out/conv_shapes.txt"""
import ctypes as C
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nvae_tf_b200 import runtime as R  # noqa: E402
from nvae_tf_b200.models import NVAE, Adamax, CosineDecay  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 144
    reps = 10
    torch.cuda.set_device(0)
    m = NVAE(**bench.mirror_kwargs(B), training=True, seed=1)
    m.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 1000)))
    rt = m.rt
    seen = OrderedDict()
    orig = R.conv2d

    def spy(rt_, x, conv, x2=None, residual=None, shift=(0, 0), out=None, y_off=0, pre=(0.0, 0.0)):
        k = conv.kernel
        key = (x.shape, x2.shape[-1] if x2 is not None else 0, k.shape, conv.stride, shift,
               out.shape[-1] if out is not None else 0, y_off)
        if key not in seen:
            seen[key] = [0, conv]
        seen[key][0] += 1
        return orig(rt_, x, conv, x2=x2, residual=residual, shift=shift, out=out, y_off=y_off, pre=pre)
    R.conv2d = spy
    x = torch.as_tensor(bench.synthetic_images(B, 1)).cuda()
    m.steps = 20000
    m.train_step(x)
    R.conv2d = orig
    torch.cuda.synchronize()
    rows = []
    ws, wsb = rt.workspace(512 << 20)
    for key, (count, conv) in seen.items():
        xs, cin2, ks, stride, shift, y_ld, y_off = key
        d = R.conv_desc(rt, xs, cin2, ks, stride, shift, y_ld, y_off)
        tc = [rt.lib._nvae_conv2d_uses_tensor_cores(C.byref(d), i) for i in range(3)]
        N, H, W, Cin = xs
        xt = torch.randn(N, H, W, Cin, device="cuda")
        x2t = torch.randn(N, H, W, cin2, device="cuda") if cin2 else None
        ld = y_ld if y_ld else ks[3]
        yt = torch.randn(d.N, d.Ho, d.Wo, ld, device="cuda")
        dx = torch.empty_like(xt)
        dx2 = torch.empty_like(x2t) if cin2 else None
        fns = (
            lambda: rt.lib.conv2d_fwd(C.byref(d), xt.data_ptr(), x2t.data_ptr() if cin2 else None, conv.kernel.ptr(),
                                      conv.packed_fwd(), None, None, yt.data_ptr(), ws, wsb, rt.stream),
            lambda: rt.lib.conv2d_dgrad(C.byref(d), yt.data_ptr(), conv.kernel.ptr(), conv.packed_dgrad(),
                                        dx.data_ptr(), dx2.data_ptr() if cin2 else None, 0, ws, wsb, rt.stream),
            lambda: rt.lib.conv2d_wgrad(C.byref(d), xt.data_ptr(), x2t.data_ptr() if cin2 else None, yt.data_ptr(),
                                        conv.kernel.gptr(), None, ws, wsb, rt.stream))
        us = []
        for fn in fns:
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.graph(g, stream=s):
                for _ in range(reps):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us.append(e0.elapsed_time(e1) * 1e3 / reps)
        fl = 2.0 * d.N * d.Ho * d.Wo * ks[3] * ks[0] * ks[1] * ks[2]
        rows.append((count * sum(us), count, key, tc, us, fl))
    rows.sort(key=lambda r: -r[0])
    tot = sum(r[0] for r in rows)
    print(f"batch {B}: {len(rows)} distinct conv shapes, {sum(r[1] for r in rows)} conv layers, "
          f"{tot / 1e3:.2f} ms fwd+dgrad+wgrad per step")
    print("   ms/step  count  shape [N,H,W,Cin]+Cin2 -> k RxS Cout /stride   tc  fwd us (TF/s)  dgrad us  wgrad us")
    for t, count, key, tc, us, fl in rows:
        xs, cin2, ks, stride, shift, y_ld, y_off = key
        print(f"{t / 1e3:9.3f}  {count:5d}  {list(xs)}+{cin2} -> {ks[0]}x{ks[1]} {ks[3]} /{stride} "
              f"{'shift' if shift != (0, 0) else ''}  {tc}  {us[0]:8.1f} ({fl / us[0] / 1e6:6.1f})  {us[1]:8.1f} "
              f"({fl / us[1] / 1e6:6.1f})  {us[2]:8.1f} ({fl / us[2] / 1e6:6.1f})")


if __name__ == "__main__":
    main()
