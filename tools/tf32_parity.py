#!/usr/bin/env python
"""Development probe: whole-step parity of the full MNIST-config NVAE (train.py defaults) against the
float64 oracle, in fp32 (CUDA-core) and TF32 (tcgen05) arithmetic.  Prints the worst per-tensor errors.
usage (GPU box): python tools/tf32_parity.py [batch] [modes...]   modes: fp32 tf32x3 tf32:dy tf32:all tf32:none
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from nvae_tf_b200 import _lib  # noqa: E402
from nvae_tf_b200.models import NVAE, Adamax, CosineDecay  # noqa: E402
from oracle import nvae_oracle as O  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    modes = sys.argv[2:] or ["fp32", "tf32:dy"]
    cfg = O.NVAEConfig()
    steps = 30000  # beta ~ 0.6: KL contributes, balancing active
    params, trainable, bnl, s = O.build_params(cfg, seed=3, jitter=0.05)
    params = {k: v.astype(np.float32).astype(np.float64) for k, v in params.items()}
    x = O.make_images(cfg, batch, seed=3).numpy()
    eps = [e.numpy().astype(np.float32).astype(np.float64) for e in O.make_eps(s, batch, seed=3)]
    t0 = time.time()
    torch.set_num_threads(os.cpu_count())
    losses, grads, c, _ = H.run_oracle_step(cfg, params, trainable, bnl, s, x, eps, steps, True)
    print(f"oracle step (fp64, batch {batch}): {time.time() - t0:.1f}s  loss {float(losses['loss']):.6f}", flush=True)
    for mode in modes:
        prec = {"fp32": _lib.NVAE_PREC_FP32, "tf32x3": _lib.NVAE_PREC_TF32X3}.get(mode, _lib.NVAE_PREC_TF32)
        if ":" in mode:
            os.environ["NVAE_TF32_ROUND"] = mode.split(":")[1]
        m = NVAE(**H.mirror_kwargs(cfg, batch), training=True, precision=prec)
        m.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 1000)))
        m.rt.load_named(params)
        m.rt.inject_eps(eps)
        m.steps = steps
        out = m.train_step(x, apply_gradients=False)
        torch.cuda.synchronize()
        loss = float(out["loss"].item())
        got = m.rt.named_grads()
        gmax = max(float(np.abs(g).max()) for g in grads.values())
        errs = []
        for n, w in grads.items():
            if np.abs(w).max() <= 1e-9 * gmax:
                continue
            errs.append((H.max_rel_err(got[n], w, 1e-4 * gmax), n))
        errs.sort(reverse=True)
        kl = H.max_rel_err(m.decoder.sampler.kl_all.cpu().numpy(), losses["kl_all"])
        rec = H.max_rel_err(out["reconstruction_loss"].cpu().numpy(), losses["reconstruction_loss"])
        print(f"[{mode}] loss {loss:.6f} rel {abs(loss - float(losses['loss'])) / abs(float(losses['loss'])):.2e} "
              f"recon {rec:.2e} kl_all {kl:.2e}; grads: worst {errs[0][0]:.2e} ({errs[0][1]}), "
              f"median {errs[len(errs) // 2][0]:.2e}, >1e-3: {sum(e > 1e-3 for e, _ in errs)}/{len(errs)}", flush=True)
        for e, n in errs[:8]:
            print(f"      {e:.2e} {n}")
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
