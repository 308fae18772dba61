#!/usr/bin/env python
"""Per-kernel roofline table (BASELINE metric: "per-kernel % of tensor/HBM roofline") and the residual-cell
micro-benchmark of BASELINE configs[1] (NHWC batch 256 at 14x14x64 and 7x7x128, plus the model's own shapes).

Bandwidth-bound kernels are timed on tensors larger than the 126 MB L2 (or with an L2 flush between repetitions)
with CUDA events around the replay of a one-call CUDA graph (the step replays the same launches from a graph, so the CPU
launch gaps between the two or three kernels of an operator are not part of it); achieved = ALGORITHMIC bytes (every logically required
tensor read/written once) / time; peak = MEASURED_PEAKS.json hbm_gbs.
usage (GPU box): python tools/kernel_roofline.py > gpurun_out/kernel_roofline.md
"""
import ctypes as C
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvae_tf_b200 import _lib  # noqa: E402
from nvae_tf_b200 import runtime as R  # noqa: E402
from nvae_tf_b200.decoder import GenerativeResidualCell  # noqa: E402
from nvae_tf_b200.encoder import EncodingResidualCell  # noqa: E402
from nvae_tf_b200.layers import BatchNormalization, DepthwiseConv2D  # noqa: E402
from nvae_tf_b200.common import SqueezeExcitation  # noqa: E402
from nvae_tf_b200.runtime import DeviceTensor  # noqa: E402

try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    PEAKS = {}
HBM = PEAKS.get("hbm_gbs", 6650.0)
FLUSH = None


def timed(fn, reps=5, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def graph_timed(fn, iters=10):
    """Device time of one fn() inside a CUDA graph (iters back-to-back calls per replay): no CPU launch overhead."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for _ in range(iters):
            fn()
    return timed(g.replay, reps=7, flush=False) / iters


def timed_op(fn, reps=5):
    """One call of a (possibly multi-launch) operator, captured into a CUDA graph so that the CPU launch overhead between
    its kernels does not count -- the training step replays these same launches from a graph -- and replayed with the L2
    flushed before every repetition."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        fn()
    return timed(g.replay, reps=reps, flush=True)


def row(name, shape, nbytes, us):
    gbs = nbytes / us / 1e3
    print(f"| `{name}` | {shape} | {nbytes / 1e6:.1f} | {us:.1f} | {gbs:.0f} | {100 * gbs / HBM:.0f}% |")


def membound():
    print(f"## Bandwidth-bound kernels (peak = {HBM:.0f} GB/s, MEASURED_PEAKS.json hbm_gbs)\n")
    print("| kernel | tensor | alg. MB | us | GB/s | of HBM peak |\n|---|---|---:|---:|---:|---:|")
    rt = R.Runtime(seed=1)
    with rt:
        shapes = [(256, 14, 14, 384), (144, 32, 32, 192), (144, 8, 8, 768)]  # decoder-cell hidden tensors (config 2, model)
        bns = {c: BatchNormalization(momentum=0.05, epsilon=1e-5, channels=c, name=f"bn{c}") for c in {s[3] for s in shapes}}
        dws = {c: DepthwiseConv2D((5, 5), padding="same", in_channels=c, name=f"dw{c}") for c in {s[3] for s in shapes}}
        ses = {c: SqueezeExcitation(channels=c, name=f"se{c}") for c in (64, 128)}
        rt.finalize()
        lib = rt.lib
        for shp in shapes:
            N, H, W, Cc = shp
            rows = N * H * W
            n = rows * Cc
            x = torch.randn(*shp, device="cuda")
            y = torch.empty_like(x)
            dy = torch.randn_like(x)
            dx = torch.empty_like(x)
            bn, dw = bns[Cc], dws[Cc]
            stat = torch.empty(4, Cc, device="cuda")
            ws, wsb = rt.workspace(max(lib._nvae_bn_ws_bytes(rows, Cc), lib._nvae_dwconv5x5_bwd_filter_ws_bytes(N, H, W, Cc)))
            tag = f"[{N},{H},{W},{Cc}]"
            us = timed_op(lambda: lib.bn_stats(x.data_ptr(), rows, Cc, bn.gamma.ptr(), bn.beta.ptr(), bn.moving_mean.ptr(),
                                            bn.moving_variance.ptr(), 1, 0.05, 1e-5, stat.data_ptr(), ws, wsb, rt.stream))
            row("bn_stats + bn_finalize", tag, 4 * n, us)
            us = timed_op(lambda: lib.bn_act_fwd(x.data_ptr(), rows, Cc, stat.data_ptr(), 1, 0, 0, 0, y.data_ptr(), rt.stream))
            row("bn_act_fwd (BN-apply + swish)", tag, 8 * n, us)
            us = timed_op(lambda: lib.bn_act_bwd(dy.data_ptr(), x.data_ptr(), rows, Cc, stat.data_ptr(), 1, 0, 0, 1, None, 0.0,
                                              0, dx.data_ptr(), bn.gamma.gptr(), bn.beta.gptr(), ws, wsb, rt.stream))
            row("bn_act_bwd (reduce + finalize + apply)", tag, 20 * n, us)
            if H <= 14:
                us = timed_op(lambda: lib.dwconv5x5_fwd(x.data_ptr(), stat.data_ptr(), 1, N, H, W, Cc, dw.depthwise_kernel.ptr(),
                                                     dw.bias.ptr(), y.data_ptr(), rt.stream))
                row("dwconv5x5_fwd (+BN-apply+swish on load)", tag, 8 * n, us)
                us = timed_op(lambda: lib.dwconv5x5_bwd_data(dy.data_ptr(), N, H, W, Cc, dw.depthwise_kernel.ptr(),
                                                          dx.data_ptr(), rt.stream))
                row("dwconv5x5_bwd_data", tag, 8 * n, us)
                us = timed_op(lambda: lib.dwconv5x5_bwd_filter(x.data_ptr(), stat.data_ptr(), 1, dy.data_ptr(), N, H, W, Cc,
                                                            dw.depthwise_kernel.gptr(), dw.bias.gptr(), ws, wsb, rt.stream))
                row("dwconv5x5_bwd_filter", tag, 8 * n, us)
            del x, y, dy, dx
        small_rows = []
        for shp in [(144, 4, 4, 256), (144, 4, 4, 1536), (144, 8, 8, 128), (144, 8, 8, 768), (144, 16, 16, 64)]:
            # the model's own L2-resident BN tensors: launch-/latency-bound, timed WITHOUT an L2 flush
            N, H, W, Cc = shp
            rows, n = N * H * W, N * H * W * Cc
            x, dy = torch.randn(*shp, device="cuda"), torch.randn(*shp, device="cuda")
            y, dx = torch.empty_like(x), torch.empty_like(x)
            g, b, mm, mv = (torch.ones(Cc, device="cuda") for _ in range(4))
            dg, db = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
            stat = torch.empty(4, Cc, device="cuda")
            ws, wsb = rt.workspace(lib._nvae_bn_ws_bytes(rows, Cc))
            def call(f):  # launch on the capturing stream
                return lambda: f(torch.cuda.current_stream().cuda_stream)
            uf = graph_timed(call(lambda st: lib.bn_fwd(x.data_ptr(), rows, Cc, g.data_ptr(), b.data_ptr(), mm.data_ptr(),
                                                        mv.data_ptr(), 1, 0.05, 1e-5, stat.data_ptr(), 1, 0, 0, 0,
                                                        y.data_ptr(), ws, wsb, st)))
            ub = graph_timed(call(lambda st: lib.bn_act_bwd(dy.data_ptr(), x.data_ptr(), rows, Cc, stat.data_ptr(), 1, 0, 0,
                                                            1, None, 0.0, 0, dx.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                                            ws, wsb, st)))
            small_rows.append((f"[{N},{H},{W},{Cc}]", 4 * n / 1e6, uf, ub))
        for shp in [(256, 14, 14, 64), (256, 7, 7, 128), (144, 32, 32, 32)]:  # cell outputs: SE + residual merge
            N, H, W, Cc = shp
            if Cc not in ses:
                continue
            se = ses[Cc]
            hid = se.dense1.units
            n = N * H * W * Cc
            t, xr = torch.randn(*shp, device="cuda"), torch.randn(*shp, device="cuda")
            y, dy = torch.empty_like(t), torch.randn_like(t)
            dt, dxr = torch.empty_like(t), torch.empty_like(t)
            pooled, hidden, gate = torch.empty(N, Cc, device="cuda"), torch.empty(N, hid, device="cuda"), torch.empty(N, Cc, device="cuda")
            ws, wsb = rt.workspace(lib._nvae_se_bwd_ws_bytes(N, Cc, hid))
            tag = f"[{N},{H},{W},{Cc}]"
            us = timed_op(lambda: lib.se_fwd(t.data_ptr(), None, xr.data_ptr(), N, H * W, Cc, hid, se.dense1.kernel.ptr(),
                                          se.dense1.bias.ptr(), se.dense2.kernel.ptr(), se.dense2.bias.ptr(), 0.1, 1.0,
                                          pooled.data_ptr(), hidden.data_ptr(), gate.data_ptr(), y.data_ptr(), rt.stream))
            row("se_fwd (pool, FCs, scale + residual)", tag, 12 * n, us)
            us = timed_op(lambda: lib.se_bwd(dy.data_ptr(), t.data_ptr(), None, N, H * W, Cc, hid, se.dense1.kernel.ptr(),
                                          se.dense2.kernel.ptr(), pooled.data_ptr(), hidden.data_ptr(), gate.data_ptr(), 0.1,
                                          1.0, dt.data_ptr(), dxr.data_ptr(), 0, se.dense1.kernel.gptr(),
                                          se.dense1.bias.gptr(), se.dense2.kernel.gptr(), se.dense2.bias.gptr(), ws, wsb,
                                          rt.stream))
            row("se_bwd", tag, 16 * n, us)
        # latent math: one group at the 8x8 scale, batch 1024 (sampling config) so it exceeds L2-trivial sizes
        B, HW, L = 4096, 64, 20
        enc, dec, eps = (torch.randn(B, HW, 2 * L, device="cuda") for _ in range(2)) if False else (None, None, None)
        enc = torch.randn(B, HW, 2 * L, device="cuda")
        dec = torch.randn(B, HW, 2 * L, device="cuda")
        eps = torch.randn(B, HW, L, device="cuda")
        z, kl = torch.empty(B, HW, L, device="cuda"), torch.empty(B, device="cuda")
        dist = torch.empty(4, B, HW, L, device="cuda")
        us = timed_op(lambda: lib.latent_fwd(enc.data_ptr(), dec.data_ptr(), eps.data_ptr(), B, HW, L, z.data_ptr(), kl.data_ptr(),
                                          None, None, dist.data_ptr(), rt.stream))
        row("latent_fwd (params, sample, KL, dist)", f"[{B},8,8,20]", 4 * B * HW * L * (2 + 2 + 1 + 1 + 4), us)
        dz, klw = torch.randn_like(z), torch.full((1,), 0.01, device="cuda")
        de, dd = torch.empty_like(enc), torch.empty_like(dec)
        us = timed_op(lambda: lib.latent_bwd(enc.data_ptr(), dec.data_ptr(), eps.data_ptr(), dz.data_ptr(), klw.data_ptr(), B, HW,
                                          L, de.data_ptr(), dd.data_ptr(), rt.stream))
        row("latent_bwd", f"[{B},8,8,20]", 4 * B * HW * L * (2 + 2 + 1 + 1 + 2 + 2), us)
        n = 40_128_896
        pbuf, g, m, v = (torch.randn(n, device="cuda") for _ in range(4))
        hyper = torch.tensor([0.5, 1e-3, 1e-3, 1, 0, 0, 0, 0], device="cuda")
        us = timed_op(lambda: lib.adamax(pbuf.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, hyper.data_ptr(), 0.9,
                                      0.999, 1e-7, 1.0, rt.stream))
        row("adamax (whole parameter arena)", "[40.1M]", 28 * n, us)
        fused = os.environ.get("NVAE_BN_FUSED", "1") != "0"
        print(f"\n## BatchNorm at the model's own L2-resident shapes (inside a CUDA graph, no L2 flush; {'one cluster launch' if fused else 'split kernels, NVAE_BN_FUSED=0'})\n")
        print("| tensor | MB | fwd: stats + apply + swish, us | bwd: reduce + dgamma/dbeta + dx, us |\n|---|---:|---:|---:|")
        for tag, mb, uf, ub in small_rows:
            print(f"| {tag} | {mb:.1f} | {uf:.1f} | {ub:.1f} |")


def cells():
    print("\n## Residual cells (BASELINE configs[1]): forward / forward+backward, 3xTF32, training-mode BN, SN packed\n")
    print("| cell | x | fwd us | fwd+bwd us | alg. GFLOP fwd+bwd | TFLOP/s (fwd+bwd) |\n|---|---|---:|---:|---:|---:|")
    for kind, shp in [("enc", (256, 14, 14, 64)), ("enc", (256, 7, 7, 128)), ("enc", (144, 8, 8, 128)), ("enc", (144, 4, 4, 256)),
                      ("dec", (256, 14, 14, 64)), ("dec", (256, 7, 7, 128)), ("dec", (144, 8, 8, 128)), ("dec", (144, 4, 4, 256))]:
        rt = R.Runtime(seed=1)
        with rt:
            N, H, W, Cc = shp
            cell = (EncodingResidualCell if kind == "enc" else GenerativeResidualCell)(Cc, name="cell")
            rt.finalize()
            rt.spectral_normalize_all(power_iter=True)
            rt.sn_done = True
            x = DeviceTensor(torch.randn(*shp, device="cuda"))
            dy = torch.randn(*shp, device="cuda")

            def fwd():
                return cell(x, training=True)

            def fwdbwd():
                x.grad = None
                with rt.gradient_tape() as tape:
                    y = cell(x, training=True)
                y.grad = dy
                rt.backward(tape)
            for f in (fwd, fwdbwd):
                f()
            torch.cuda.synchronize()
            res = []
            for f in (fwd, fwdbwd):
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.graph(g, stream=s):
                    f()
                res.append(timed(g.replay, reps=5, flush=True))
            px = N * H * W
            macs = px * (2 * 9 * Cc * Cc) if kind == "enc" else px * (2 * 6 * Cc * Cc)
            gf = 6.0 * macs / 1e9  # fwd + dgrad + wgrad, 2 flops per MAC
            print(f"| {'EncodingResidualCell' if kind == 'enc' else 'GenerativeResidualCell'} | {list(shp)} | {res[0]:.0f} | "
                  f"{res[1]:.0f} | {gf:.1f} | {gf / res[1] / 1e-3:.1f} |")


if __name__ == "__main__":
    torch.cuda.set_device(0)
    print("# Per-kernel roofline and residual-cell micro-benchmark (B200, measured by tools/kernel_roofline.py)\n")
    membound()
    if "--no-cells" not in sys.argv:
        cells()
