#!/usr/bin/env python
"""One launch (after one warm-up, L2 flushed in between) of each kernel the ncu evidence under profiles/ is about:
the dominant 3xFP16 convolution (forward / backward-data / backward-filter of postprocess cbs2, 5x5 384->384 @16x16, batch 144)
and the bandwidth kernels at SURVEY 8d's shapes (depthwise 5x5, squeeze-excitation, BatchNorm statistics / backward).
Run plain first, then under ncu (B200_PROFILING.md):
    python tools/profile_kernels.py && ncu --set full --clock-control none --import-source on \\
        -k regex:'conv_tc_kernel|dwconv5x5|se_|bn_' -o gpurun_out/r02_kernels python tools/profile_kernels.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvae_tf_b200 import _lib  # noqa: E402
from nvae_tf_b200 import runtime as R  # noqa: E402
from nvae_tf_b200.common import SqueezeExcitation  # noqa: E402
from nvae_tf_b200.layers import BatchNormalization, Conv2D, DepthwiseConv2D  # noqa: E402


def main():
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with R.Runtime(seed=1, precision=_lib.NVAE_PREC_TF32X3) as rt:
        conv = Conv2D(384, (5, 5), padding="same", in_channels=384, name="cbs2")
        bn384 = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=384, name="bn384")
        bn192 = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=192, name="bn192")
        dw = DepthwiseConv2D((5, 5), padding="same", in_channels=384, name="dw")
        se = SqueezeExcitation(channels=64, name="se")
        rt.finalize()
        lib = rt.lib
        rt.pack_plain(conv)

        def once(fn):
            for _ in range(2):  # warm-up, then the launch ncu's -s/-c window should pick (every second one)
                flush.zero_()
                fn()
            torch.cuda.synchronize()

        # ---- dominant GEMM ------------------------------------------------------------------------------------------
        x = torch.randn(144, 16, 16, 384, device="cuda")
        dy = torch.randn(144, 16, 16, 384, device="cuda") * 1e-3
        y, dx = torch.empty_like(x), torch.empty_like(x)
        d = R.conv_desc(rt, tuple(x.shape), 0, conv.kernel.shape, 1)
        ws, wsb = rt.workspace(max(lib._nvae_conv2d_ws_bytes(C.byref(d), i) for i in range(3)))
        once(lambda: lib.conv2d_fwd(C.byref(d), x.data_ptr(), None, conv.kernel.ptr(), conv.packed_fwd(), None, None,
                                    y.data_ptr(), ws, wsb, rt.stream))
        once(lambda: lib.conv2d_dgrad(C.byref(d), dy.data_ptr(), conv.kernel.ptr(), conv.packed_dgrad(), dx.data_ptr(), None, 0,
                                      ws, wsb, rt.stream))
        once(lambda: lib.conv2d_wgrad(C.byref(d), x.data_ptr(), None, dy.data_ptr(), conv.kernel.gptr(), None, ws, wsb, rt.stream))
        del x, dy, y, dx
        # ---- depthwise 5x5 at [256,14,14,384] (BASELINE configs[1] decoder-cell hidden tensor) ---------------------------
        N, H, W, Cc = 256, 14, 14, 384
        rows = N * H * W
        x = torch.randn(N, H, W, Cc, device="cuda")
        dy = torch.randn_like(x)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        stat = torch.empty(4, Cc, device="cuda")
        ws, wsb = rt.workspace(max(lib._nvae_bn_ws_bytes(rows, Cc), lib._nvae_dwconv5x5_bwd_filter_ws_bytes(N, H, W, Cc)))
        once(lambda: lib.bn_stats(x.data_ptr(), rows, Cc, bn384.gamma.ptr(), bn384.beta.ptr(), bn384.moving_mean.ptr(),
                                  bn384.moving_variance.ptr(), 1, 0.05, 1e-5, stat.data_ptr(), ws, wsb, rt.stream))
        once(lambda: lib.dwconv5x5_fwd(x.data_ptr(), stat.data_ptr(), 1, N, H, W, Cc, dw.depthwise_kernel.ptr(), dw.bias.ptr(),
                                       y.data_ptr(), rt.stream))
        once(lambda: lib.dwconv5x5_bwd_data(dy.data_ptr(), N, H, W, Cc, dw.depthwise_kernel.ptr(), dx.data_ptr(), rt.stream))
        once(lambda: lib.dwconv5x5_bwd_filter(x.data_ptr(), stat.data_ptr(), 1, dy.data_ptr(), N, H, W, Cc,
                                              dw.depthwise_kernel.gptr(), dw.bias.gptr(), ws, wsb, rt.stream))
        del x, dy, y, dx
        # ---- BatchNorm on the largest tensor of the step, [144,32,32,192] (113 MB) ------------------------------------
        N, H, W, Cc = 144, 32, 32, 192
        rows = N * H * W
        x = torch.randn(N, H, W, Cc, device="cuda")
        dy = torch.randn_like(x)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        stat = torch.empty(4, Cc, device="cuda")
        ws, wsb = rt.workspace(lib._nvae_bn_ws_bytes(rows, Cc))
        once(lambda: lib.bn_stats(x.data_ptr(), rows, Cc, bn192.gamma.ptr(), bn192.beta.ptr(), bn192.moving_mean.ptr(),
                                  bn192.moving_variance.ptr(), 1, 0.05, 1e-5, stat.data_ptr(), ws, wsb, rt.stream))
        once(lambda: lib.bn_act_fwd(x.data_ptr(), rows, Cc, stat.data_ptr(), 1, 0, 0, 0, y.data_ptr(), rt.stream))
        once(lambda: lib.bn_act_bwd(dy.data_ptr(), x.data_ptr(), rows, Cc, stat.data_ptr(), 1, 0, 0, 1, None, 0.0, 0,
                                    dx.data_ptr(), bn192.gamma.gptr(), bn192.beta.gptr(), ws, wsb, rt.stream))
        del x, dy, y, dx
        # ---- squeeze-excitation + residual merge at [256,14,14,64] -------------------------------------------------
        N, H, W, Cc = 256, 14, 14, 64
        hid = se.dense1.units
        t, xr = torch.randn(N, H, W, Cc, device="cuda"), torch.randn(N, H, W, Cc, device="cuda")
        y, dy = torch.empty_like(t), torch.randn_like(t)
        dt, dxr = torch.empty_like(t), torch.empty_like(t)
        pooled, hidden, gate = (torch.empty(N, Cc, device="cuda"), torch.empty(N, hid, device="cuda"),
                                torch.empty(N, Cc, device="cuda"))
        ws, wsb = rt.workspace(lib._nvae_se_bwd_ws_bytes(N, Cc, hid))
        once(lambda: lib.se_fwd(t.data_ptr(), None, xr.data_ptr(), N, H * W, Cc, hid, se.dense1.kernel.ptr(),
                                se.dense1.bias.ptr(), se.dense2.kernel.ptr(), se.dense2.bias.ptr(), 0.1, 1.0,
                                pooled.data_ptr(), hidden.data_ptr(), gate.data_ptr(), y.data_ptr(), rt.stream))
        once(lambda: lib.se_bwd(dy.data_ptr(), t.data_ptr(), None, N, H * W, Cc, hid, se.dense1.kernel.ptr(),
                                se.dense2.kernel.ptr(), pooled.data_ptr(), hidden.data_ptr(), gate.data_ptr(), 0.1, 1.0,
                                dt.data_ptr(), dxr.data_ptr(), 0, se.dense1.kernel.gptr(), se.dense1.bias.gptr(),
                                se.dense2.kernel.gptr(), se.dense2.bias.gptr(), ws, wsb, rt.stream))
    print("profile_kernels: done")


if __name__ == "__main__":
    main()
