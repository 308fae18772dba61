#!/usr/bin/env python
"""bench.py -- NVAE train images/s on N B200s + roofline of the dominant kernel + CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # the reference's algorithm on the host CPU

Workload (BASELINE.json configs[2]): full MNIST-config NVAE train step -- train.py defaults, 40.1 M
parameters, all 15 latent groups, KL + reconstruction + BN-gamma losses, spectral norm, Adamax --
batch 144 per GPU on synthetic binarised 28x28 images zero-padded to 32x32.  A "step" is one whole
training step (fill, SN power iteration, forward, losses, backward, [all-reduce], Adamax).

`value`  : images/s with the batch already resident in HBM (CUDA-graph replay), CUDA-event timed, max over ranks.
`e2e`    : images/s through the public host-batch API (NVAE.make_train_function): per step a pinned
           host->device copy of the batch and a device->host read of the four losses.
`roofline`: the dominant kernel (largest forward convolution, postprocess.py:74-76 shape) timed alone with
           CUDA events on the launching stream, L2 flushed between repetitions.
`cpu_baseline`: the CPU oracle port (torch fp32, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "nvae_train_images_per_sec"
UNIT = "images/s"
BATCH = 144
FLOP_PER_IMAGE = 38.87e9  # 6 x 6.478 GMAC: fwd + dgrad + wgrad


def default_cfg():
    from oracle import nvae_oracle as O  # config dataclass only; the oracle is never on the GPU path
    return O.NVAEConfig()


def mirror_kwargs(batch):
    return dict(n_encoder_channels=32, n_decoder_channels=32, res_cells_per_group=1, n_preprocess_blocks=2,
                n_preprocess_cells=3, n_latent_per_group=20, n_latent_scales=2, n_groups_per_scale=[5, 10],
                n_postprocess_blocks=2, n_post_process_cells=3, sr_lambda=0.01, scale_factor=2, total_epochs=400,
                n_total_iterations=417 * 400, step_based_warmup=True, input_shape=[batch, 32, 32, 1])


def synthetic_images(batch, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    x = (rng.random((batch, 28, 28, 1)) < 0.13).astype(np.float32)
    return np.pad(x, ((0, 0), (2, 2), (2, 2), (0, 0)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm (TensorFlow itself is not installable here)
# ------------------------------------------------------------------------------------------
def cpu_step_fn(batch, threads):
    import numpy as np
    import torch
    import helpers as H  # noqa: F401
    from oracle import nvae_oracle as O
    torch.set_num_threads(threads)
    cfg = O.NVAEConfig()
    params, trainable, bnl, s = O.build_params(cfg, seed=1)
    pt = O.to_torch(params, trainable, dtype=torch.float32)
    x = torch.as_tensor(synthetic_images(batch, 1))
    eps = [e.float() for e in O.make_eps(s, batch, seed=1)]

    def step(i):
        out, c = O.train_step_loss(cfg, s, pt, bnl, x, eps, steps=20000 + i, training=True)
        O.grads_wrt_trainables(out["loss"], c, trainable)
        return float(out["loss"].detach())
    return step


def cpu_baseline(sample_batch, budget_s=12.0):
    threads = os.cpu_count() or 1
    step = cpu_step_fn(sample_batch, threads)
    step(0)  # warm-up (thread pools, oneDNN primitive caches)
    n, t0 = 0, time.perf_counter()
    while n < 2 or (time.perf_counter() - t0 < budget_s and n < 64):
        n += 1
        step(n)
    dt = time.perf_counter() - t0
    return {"value": n * sample_batch / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"1 warm-up + {n} timed full train steps (fwd+bwd, no optimizer) of the oracle port in torch-CPU "
                      f"fp32 at batch {sample_batch} of the same config (~{budget_s:.0f} s of CPU work); TensorFlow (the "
                      f"reference's runtime) is not installable in this image", "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_batch
    step = cpu_step_fn(sample, threads)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "full MNIST-config NVAE train step (train.py defaults, 40.1M params)",
                       "batch_per_step": sample, "note": "bounded CPU sample of the batch-144 workload"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} timed steps at batch {sample}, oracle port, torch-CPU fp32"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# roofline of the dominant kernel, measured live
# ------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant 3xFP16 launch (profiles/r01i_ncu_conv_f16_5x5_384.md)
DOMINANT_TRAFFIC_F16 = 104.1e6


def measure_dominant_kernel(model, reps=5):
    """Largest forward conv of the step: postprocess cbs2, 5x5, 384->384 at 16x16, batch 144
    (GEMM M=36864, N=384, K=9600; 2MNK = 271.8 GFLOP per launch).  Timed alone, L2 flushed between reps."""
    import torch
    from nvae_tf_b200 import runtime as R
    from nvae_tf_b200.runtime import DeviceTensor
    rt = model.rt
    node = model.postprocess.cells[1].node  # block 0, non-upscaling cell: 64 ch -> hidden 384
    conv = node.cbs2.conv.layer
    k = conv.kernel
    x = DeviceTensor(torch.randn(BATCH, 16, 16, k.shape[2], device=rt.device), needs_grad=False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=rt.device)
    out = {}
    M, N, K = BATCH * 16 * 16, k.shape[3], k.shape[0] * k.shape[1] * k.shape[2]
    flops = 2.0 * M * N * K
    times = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = R.conv2d(rt, x, conv)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    ms = statistics.mean(times)
    out["conv_fwd"] = {"ms": ms, "flops": flops, "bytes": 4.0 * (x.data.numel() + y.data.numel() + k.size)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default=os.environ.get("NVAE_PRECISION", "auto"), choices=["auto", "fp32", "tf32", "tf32x3"])
    ap.add_argument("--cpu-batch", type=int, default=16, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from nvae_tf_b200 import _lib, parallel
    from nvae_tf_b200.models import NVAE, Adamax, CosineDecay

    # rank 0 prints ONE JSON line on stdout: NCCL's version banner / warnings (stdout at NCCL_DEBUG=VERSION or WARN, which
    # this image sets) go to stderr instead
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    world = parallel.init_from_env()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    precision = {"fp32": _lib.NVAE_PREC_FP32, "tf32": _lib.NVAE_PREC_TF32, "tf32x3": _lib.NVAE_PREC_TF32X3,
                 "auto": _lib.default_precision()}[args.precision]
    B = args.batch
    model = NVAE(**mirror_kwargs(B), training=True, precision=precision, seed=1)
    model.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 400 * 417)), run_eagerly=True)
    if world > 1:
        parallel.broadcast_parameters([model.rt.params, model.rt.state])
    model.steps = 20000  # inside the KL warm-up: beta < 1, balancing active (the common case of the first 30 %)
    train_fn = model.make_train_function((B, 32, 32, 1))
    static_in, replay = train_fn.static_in, train_fn.replay
    x = torch.as_tensor(synthetic_images(B, 1 + rank)).to(model.rt.device)
    static_in.copy_(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = replay()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = replay()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=model.rt.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    loss = float(out["loss"].item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the host-batch API -------------------------------------------------------------
    host_batches = [torch.as_tensor(synthetic_images(B, 100 + rank + i)).pin_memory() for i in range(4)]
    for i in range(2):
        train_fn(host_batches[i % 4])
    barrier()
    e0.record()
    for i in range(args.steps):
        res = train_fn(host_batches[i % 4])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=model.rt.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e = {"value": world * B * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": train_fn.h2d_bytes,
           "d2h_bytes_per_step": train_fn.d2h_bytes, "ms_per_step": e2e_ms / args.steps,
           "api": "NVAE.make_train_function(batch_shape)(host_batch)"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel, live ----------------------------------------------------------------
    dom = measure_dominant_kernel(model)["conv_fwd"]
    tc = precision != _lib.NVAE_PREC_FP32
    x3 = precision == _lib.NVAE_PREC_TF32X3
    # NVAE_PREC_TF32X3 runs its large GEMMs (this one included) as 3xFP16 on kind::f16 unless NVAE_F16X3=0
    f16 = x3 and os.environ.get("NVAE_F16X3", "1") != "0"
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    # kind::f16 runs at the dense bf16/fp16 rate (the measured cuBLAS figure); kind::tf32 at half of it (nominal ratio)
    peak = bf16_peak if f16 else bf16_peak / 2.0
    achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
    arith = ("3xFP16: amax-scaled fp16 hi/lo split, A via TMEM, pre-split B via TMA, two accumulators per CTA" if f16 else
             "3xTF32 split in-kernel, A via TMEM" if x3 else "single-pass TF32")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full (profiles/r01i_ncu_*.md /
                # r01h for 3xTF32); algorithmic bytes are 127.9 MB (x + w + y once)
                "traffic": DOMINANT_TRAFFIC_F16 if f16 else (113.3e6 if x3 else None),
                "mma_issue_factor": 3 if x3 else 1,
                "issued_mma_frac": (3 if x3 else 1) * achieved / peak,
                "note": ("fp32-grade products cost 3 MMAs each (hi*hi + hi*lo + lo*hi), so the ceiling of frac is 1/3; "
                         "issued_mma_frac is the tensor pipe's own utilisation against the same peak") if x3 else "",
                "kernel": ("conv_tc_kernel<false> (tcgen05 " + ("kind::f16" if f16 else "kind::tf32") +
                           " implicit GEMM, TMA-staged, " + arith + ")" +
                           (" + absmax2_kernel + f16_pack_b_kernel (operand scales / B split, inside ms_per_launch)" if f16 else "")
                           if tc else "simt_conv_kernel<FwdProb> (fp32 CUDA-core implicit GEMM)"),
                "shape": "postprocess cbs2 5x5 384->384 @16x16, batch 144: M=36864 N=384 K=9600",
                "flops_per_launch": dom["flops"], "ms_per_launch": dom["ms"],
                "peak_source": (("MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590 bf16 TFLOP/s") +
                                ("" if f16 else " / 2 for TF32")),
                "step_tensor_frac": (FLOP_PER_IMAGE * B / ((ms / args.steps) * 1e-3) / 1e12) /
                                    (peaks.get("bf16_tflops_sustained", 1400.0) / (1.0 if f16 else 2.0))}
    cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(args.cpu_batch)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("f16x3+tf32x3" if f16 else "tf32x3" if x3 else "tf32") if tc else "f32", "data": "synthetic",
            "config": {"workload": "full MNIST-config NVAE train step (train.py defaults: 40.1M params, 15 latent "
                                   "groups, KL+recon+BN-gamma loss, SN, Adamax), BASELINE configs[2]",
                       "batch_per_gpu": B, "global_batch": B * world, "image": "32x32x1",
                       "parallelism": f"dp{world}", "training_mode": "batch-stat BN + SN power iteration",
                       "l2": "per-step working set ~6 GB of activations >> 126 MB L2 (no explicit flush needed)",
                       "execution": "one CUDA graph per step" + ("" if world == 1 else " + NCCL all-reduce + Adamax")},
            "e2e": e2e, "gpu_launches": model.graph_kernels * args.steps, "kernels_per_step": model.graph_kernels,
            "launcher_calls_per_step": model.graph_launches,
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "final_loss": loss, "e2e_loss": res["loss"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
