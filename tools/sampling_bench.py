#!/usr/bin/env python
"""BASELINE configs[3]: decoder-only ancestral sampling (models.py:137-178) at temperature 0.7, batch 1024, inference-mode
BN (moving statistics), no SN.  Reports images/s and the algorithmic tensor rate (12.1 GFLOP/image).
usage (GPU box): python tools/sampling_bench.py [n_samples]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nvae_tf_b200.models import NVAE  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    torch.cuda.set_device(0)
    m = NVAE(**bench.mirror_kwargs(n), training=False, seed=1)
    for _ in range(2):
        m.sample(n_samples=n, temperature=0.7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        images, *_ = m.sample(n_samples=n, temperature=0.7)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"sample(n_samples={n}, temperature=0.7): {ms:.1f} ms, {n / ms * 1e3:.0f} images/s, "
          f"{12.1e9 * n / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic (3xTF32), images {tuple(images.shape)} "
          f"in [{float(images.min()):.3f}, {float(images.max()):.3f}]")


if __name__ == "__main__":
    main()
