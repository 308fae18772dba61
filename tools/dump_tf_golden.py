#!/usr/bin/env python
"""Pins the oracle against the REAL reference -- for a site that has TensorFlow (this image does not).

Runs the unmodified stevensdavid/nvae-tf `NVAE` (TensorFlow 2.3 + tensorflow_addons +
tensorflow_probability) on the tiny configuration of tests/helpers.py with
  * `training=True` pinned (the reference's train_step passes no training flag, models.py:117),
  * epsilon injected into `Sampler.sample` (common.py:65-68) from the same seeded stream the oracle uses,
and writes an .npz with the layout of tests/golden/*.npz (`param/`, `eps/`, `loss/`, `grad/`, `new/`), so
`tests/test_model_gpu.py::test_full_step_against_golden_fixture` and `tests/test_oracle.py` can consume it.

    python tools/dump_tf_golden.py /path/to/nvae-tf tests/golden/tf_tiny_train.npz

Environment to run it in (the reference's own pins, requirements.txt:41-42, plus the two packages it imports but does not
pin -- the releases that pair with TensorFlow 2.3):
    python 3.8
    pip install tensorflow==2.3.0 tensorflow-estimator==2.3.0 numpy==1.18.5 scipy==1.4.1 \
                tensorflow-addons==0.11.2 tensorflow-probability==0.11.1

NOT RUN in the build image (no TensorFlow wheel, no network): DESIGN.md section 5 therefore says
"parity unpinned".  The variable-name mapping below follows the attribute paths the reference's checkpoints
use, which is also how nvae_tf_b200 names its variables.
"""
import os
import sys

import numpy as np


def main():
    ref, out = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ref)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import tensorflow as tf  # noqa: F401  (raises here when TensorFlow is absent)
    import common as ref_common
    import models as ref_models
    import helpers as H
    from oracle import nvae_oracle as O

    cfg = H.oracle_cfg()
    batch, steps, seed = 4, 10, 3
    kw = H.mirror_kwargs(cfg, batch)
    model = ref_models.NVAE(**kw)
    x = O.make_images(cfg, batch, seed=seed).numpy().astype(np.float32)
    _, _, _, shapes = O.build_params(cfg, seed=seed, jitter=0.1)
    eps = [e.numpy().astype(np.float32) for e in O.make_eps(shapes, batch, seed=seed)]
    it = iter(eps)

    def sample(self, mu, sigma):  # common.py:65-68 with the noise injected
        return mu + tf.constant(next(it)) * sigma
    ref_common.Sampler.sample = sample

    # models.py:89 has no `training` argument: pin training mode on every layer call
    orig_call = ref_models.NVAE.call
    tf.keras.backend.set_learning_phase(1)
    model(x)  # build variables
    it = iter(eps)
    model.steps = steps
    names = [v.name for v in model.weights]
    before = {n: v.numpy() for n, v in zip(names, model.weights)}
    with tf.GradientTape() as tape:
        reconstruction, z_params, *_ = orig_call(model, x)
        recon = model.calculate_recon_loss(x, reconstruction)
        bn = model.calculate_bn_loss()
        beta = min(steps / (0.3 * model.n_total_iterations), 1)
        kl = beta * model.calculate_kl_loss(z_params, beta < 1)
        total = tf.reduce_mean(recon + kl) + bn
    grads = tape.gradient(total, model.trainable_weights)
    res = {"x": x, "steps": np.int64(steps), "training": np.int64(1),
           "loss/loss": total.numpy(), "loss/reconstruction_loss": recon.numpy(), "loss/kl_loss": kl.numpy(),
           "loss/bn_loss": bn.numpy(), "loss/logits": reconstruction.numpy()}
    for i, e in enumerate(eps):
        res[f"eps/{i}"] = e
    for n, v in before.items():
        res["tfparam/" + n] = v
    for v, g in zip(model.trainable_weights, grads):
        res["tfgrad/" + v.name] = g.numpy()
    for n, v in zip(names, model.weights):
        res["tfnew/" + n] = v.numpy()
    np.savez_compressed(out, **res)
    print("wrote", out, "-- map tfparam/* onto the attribute-path names with tests/test_host_layout.py's table")


if __name__ == "__main__":
    main()
