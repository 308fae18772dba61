#!/usr/bin/env python
"""BASELINE configs[4]: CIFAR-10-shape (32x32x3) synthetic NVAE with a deeper hierarchy -- three latent scales
(8x8, 4x4, 2x2), n_groups_per_scale = [5, 10, 20] = 35 groups -- one full train step (fwd + losses + bwd + Adamax) from a
CUDA graph on ONE GPU.  The head keeps the reference's single output channel (postprocess.py:29: logits broadcast over
the 3 input channels in the Bernoulli log-likelihood, README.md:25-27).  The path shards by samples exactly like the
MNIST config (bench.py --gpus N measures that scaling); this tool reports the per-GPU rate of the deeper model.
usage (GPU box): python tools/cifar_bench.py [batch] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nvae_tf_b200.models import NVAE, Adamax, CosineDecay  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 144
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    torch.cuda.set_device(0)
    m = NVAE(n_encoder_channels=32, n_decoder_channels=32, res_cells_per_group=1, n_preprocess_blocks=2,
             n_preprocess_cells=3, n_latent_per_group=20, n_latent_scales=3, n_groups_per_scale=[5, 10, 20],
             n_postprocess_blocks=2, n_post_process_cells=3, sr_lambda=0.01, scale_factor=2, total_epochs=400,
             n_total_iterations=417 * 400, step_based_warmup=True, input_shape=[B, 32, 32, 3], training=True, seed=1)
    m.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 417 * 400)))
    m.steps = 30000  # beta inside the warm-up: KL balancing on
    rng = np.random.default_rng(1)
    x = (rng.random((B, 32, 32, 3)) < 0.13).astype(np.float32)
    static_in, replay = m.capture_train_step((B, 32, 32, 3))
    static_in.copy_(torch.as_tensor(x))
    for _ in range(3):
        out = replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_params = int(m.rt.params.numel())
    print(f"CIFAR-shape NVAE (32x32x3, groups [5,10,20] = {m.decoder.sampler.n_groups}, {n_params / 1e6:.1f} M parameters), "
          f"batch {B}, 1 GPU: {ms:.1f} ms/step, {B / ms * 1e3:.0f} images/s, {m.graph_kernels} kernels per step, "
          f"loss {float(out['loss'].item()):.2f}")


if __name__ == "__main__":
    main()
