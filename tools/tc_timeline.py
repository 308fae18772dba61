#!/usr/bin/env python
"""Development probe: per-CTA timeline inside conv_tc_kernel (needs the `make -C nvae_tf_b200/csrc dbg` build).
usage (GPU box): NVAE_LIB=nvae_tf_b200/libnvae_b200_dbg.so python tools/tc_timeline.py N H W Cin Cout k [which]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvae_tf_b200 import runtime as R  # noqa: E402
from nvae_tf_b200.layers import Conv2D  # noqa: E402


def main():
    N, H, W, Cin, Cout, k = [int(v) for v in sys.argv[1:7]]
    which = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    rt = R.Runtime(seed=1)
    with rt:
        conv = Conv2D(Cout, (k, k), padding="same", in_channels=Cin, name="c")
        rt.finalize()
        rt.pack_plain(conv)
        x = torch.randn(N, H, W, Cin, device="cuda")
        y = torch.randn(N, H, W, Cout, device="cuda")
        dx = torch.empty_like(x)
        d = R.conv_desc(rt, tuple(x.shape), 0, conv.kernel.shape, 1)
        ws, wsb = rt.workspace(256 << 20)
        fn = [lambda: rt.lib.conv2d_fwd(C.byref(d), x.data_ptr(), None, conv.kernel.ptr(), conv.packed_fwd(), None, None,
                                        y.data_ptr(), ws, wsb, rt.stream),
              lambda: rt.lib.conv2d_dgrad(C.byref(d), y.data_ptr(), conv.kernel.ptr(), conv.packed_dgrad(), dx.data_ptr(),
                                          None, 0, ws, wsb, rt.stream),
              lambda: rt.lib.conv2d_wgrad(C.byref(d), x.data_ptr(), None, y.data_ptr(), conv.kernel.gptr(), None, ws, wsb,
                                          rt.stream)][which]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"event time {e0.elapsed_time(e1) * 1e3:.1f} us (incl. fix-up)")
        if os.environ.get("TC_PROFILE"):
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
            for e in prof.events():
                if e.device_type == torch.autograd.DeviceType.CUDA:
                    print(f"   {e.device_time:8.1f} us  {e.name[:90]}")
        if not hasattr(rt.lib.dll, "nvae_debug_tc_timing"):
            return  # regular build: timing only (this script doubles as the single-kernel workload for ncu)
        buf = (C.c_ulonglong * (148 * 16))()
        rc = rt.lib.dll.nvae_debug_tc_timing(buf)
        t = np.array(buf, dtype=np.float64).reshape(148, 16)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        t = (t - t0) / 1e3
        names = ["start", "setup", "tile0", "mma_issued", "acc_done", "epi_done", "exit", "tma8 issue", "cv8 begin", "cv8 lo_empty", "cv8 full",
                 "cv8 converted", "cv8 arrived", "cv9 arrived", "mma8 conv ok", "mma9 conv ok"]
        print(f"rc={rc} event time {e0.elapsed_time(e1) * 1e3:.1f} us (incl. fix-up), {len(t)} CTAs; us since first CTA start:")
        for i, n in enumerate(names):
            print(f"  {n:11s} min {t[:, i].min():7.2f}  median {np.median(t[:, i]):7.2f}  max {t[:, i].max():7.2f}")
        if hasattr(rt.lib.dll, "nvae_debug_tc_wait"):
            wb = (C.c_longlong * 8)()
            rt.lib.dll.nvae_debug_tc_wait(wb)
            full, conv, total, n = wb[0], wb[1], wb[2], max(wb[3], 1)
            print(f"CTA 0, issuer 0, whole launch: {n} k-units, {total / n:.0f} cycles per unit, of which waiting on full[] (TMA) "
                  f"{full / n:.0f}, on conv[] (converters) {conv / n:.0f}")
        if hasattr(rt.lib.dll, "nvae_debug_tc_cycles"):
            cb = (C.c_longlong * 128)()
            rt.lib.dll.nvae_debug_tc_cycles(cb)
            cy = np.array(cb, dtype=np.int64).reshape(8, 16)
            base = cy[0, 0]
            print("CTA 0, cycles since MMA thread reached unit 8 (units 8..15):")
            print("  unit  mma:wait_conv  conv_ok  issued+commit | conv:begin  lo_empty_ok  full_ok  arrived")
            for u in range(8):
                r = cy[u] - base
                print(f"  {u + 8:4d}  {r[0]:12d} {r[1]:8d} {r[2]:14d} | {r[3]:10d} {r[4]:12d} {r[5]:8d} {r[6]:8d}")
            print("  unit  loop_top  full_ok  wait_conv  conv_ok  mma0  +4  +8  +12  commit1  commit2   (MMA thread, same origin)")
            for u in range(8):
                r = cy[u] - base
                print(f"  {u + 8:4d}  {r[13]:8d} {r[14]:8d} {r[0]:9d} {r[1]:8d} {r[8]:6d} {r[9]:5d} {r[10]:5d} {r[11]:5d} {r[2]:8d} {r[12]:8d}")


if __name__ == "__main__":
    main()
