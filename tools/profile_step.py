#!/usr/bin/env python
"""ONE replay of the captured batch-144 train step between cudaProfilerStart / Stop, for the ncu launch list
(profiles/r02_launches.md):
    python tools/profile_step.py && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \\
        --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nvae_tf_b200.models import NVAE, Adamax, CosineDecay  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 144
    torch.cuda.set_device(0)
    m = NVAE(**bench.mirror_kwargs(B), training=True, seed=1)
    m.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 400 * 417)))
    m.steps = 20000
    static_in, replay = m.capture_train_step((B, 32, 32, 1))
    static_in.copy_(torch.as_tensor(bench.synthetic_images(B, 1)))
    for _ in range(2):
        replay()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = replay()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"profile_step: {m.graph_kernels} kernels per step, loss {float(out['loss'].item()):.3f}")


if __name__ == "__main__":
    main()
