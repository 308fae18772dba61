"""Generates tests/golden/*.npz from the CPU oracle (oracle/nvae_oracle.py, float64).

PARITY UNPINNED: TensorFlow is not installable in this image, so these vectors come from the
restatement, not from the reference itself.  They pin the oracle against silent
drift and give the GPU tests fixed inputs/outputs that travel to the GPU box.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from oracle import nvae_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def model_case(name, batch, steps, training, seed):
    cfg = H.oracle_cfg()
    params, trainable, bnl, s = O.build_params(cfg, seed=seed, jitter=0.1)
    # inputs are rounded to float32 first so the fp32 device path sees exactly what the oracle saw
    params = {k: v.astype(np.float32).astype(np.float64) for k, v in params.items()}
    x = O.make_images(cfg, batch, seed=seed).numpy()
    eps = [e.numpy().astype(np.float32).astype(np.float64) for e in O.make_eps(s, batch, seed=seed)]
    losses, grads, c, record = H.run_oracle_step(cfg, params, trainable, bnl, s, x, eps, steps, training)
    keep = ["preprocess", "encoder/final", "decoder", "logits", "z/0", "z/3"]
    out = {"x": x.astype(np.float32), "steps": np.int64(steps), "training": np.int64(training)}
    for i, e in enumerate(eps):
        out[f"eps/{i}"] = e.astype(np.float32)
    for k, v in params.items():
        out["param/" + k] = v.astype(np.float32)
    for k, v in losses.items():
        out["loss/" + k] = v
    for k in keep:
        out["act/" + k] = record[k].detach().numpy()
    for k, v in grads.items():
        out["grad/" + k] = v.astype(np.float32)
    for k, v in c.new_stats.items():  # updated BN moving statistics and SN `u` (normalised kernels: sigma only)
        if k.endswith("/kernel"):
            out["sigma/" + k] = np.float64(np.abs(params[k]).max() / np.abs(v.detach().numpy()).max())
        else:
            out["new/" + k] = v.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: float(np.asarray(v).sum()) for k, v in losses.items() if k != "logits"})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    model_case("tiny_train_balanced", batch=4, steps=10, training=True, seed=3)   # beta = 1/3, KL balancing on
    model_case("tiny_infer_beta1", batch=3, steps=1000, training=False, seed=4)   # beta = 1, moving stats, no SN
