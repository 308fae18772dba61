"""Shared test helpers: tiny configs, oracle <-> mirror plumbing, error metrics."""
import numpy as np
import torch

from oracle import nvae_oracle as O

# small but structurally complete config: 2 scales, BN channels % 4 == 0, depthwise channels % 32 == 0
TINY = dict(n_encoder_channels=4, n_decoder_channels=4, res_cells_per_group=1, n_preprocess_blocks=2,
            n_preprocess_cells=2, n_latent_per_group=4, n_groups_per_scale=(2, 2), n_postprocess_blocks=2,
            n_post_process_cells=2, sr_lambda=0.01, scale_factor=2, total_epochs=2, n_total_iterations=100,
            step_based_warmup=True)


def oracle_cfg(**over):
    d = dict(TINY)
    d.update(over)
    return O.NVAEConfig(**d)


def mirror_kwargs(cfg: O.NVAEConfig, batch: int):
    return dict(n_encoder_channels=cfg.n_encoder_channels, n_decoder_channels=cfg.n_decoder_channels,
                res_cells_per_group=cfg.res_cells_per_group, n_preprocess_blocks=cfg.n_preprocess_blocks,
                n_preprocess_cells=cfg.n_preprocess_cells, n_latent_per_group=cfg.n_latent_per_group,
                n_latent_scales=len(cfg.n_groups_per_scale), n_groups_per_scale=list(cfg.n_groups_per_scale),
                n_postprocess_blocks=cfg.n_postprocess_blocks, n_post_process_cells=cfg.n_post_process_cells,
                sr_lambda=cfg.sr_lambda, scale_factor=cfg.scale_factor, total_epochs=cfg.total_epochs,
                n_total_iterations=cfg.n_total_iterations, step_based_warmup=cfg.step_based_warmup,
                input_shape=[batch, cfg.image_size, cfg.image_size, cfg.image_channels])


def max_rel_err(a, b, floor: float = 0.0) -> float:
    """Per-tensor max relative error as BASELINE's north star states it: max|a-b| / max|b|.
    `floor` bounds the denominator from below for tensors that are analytically zero (e.g. the bias
    of a conv feeding a training-mode BatchNorm has an exactly-zero gradient; only round-off remains)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = max(np.abs(b).max(), floor)
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def grad_floor(grads: dict, frac: float = 1e-4) -> float:
    """Denominator floor for gradient comparisons: `frac` of the largest gradient entry in the model."""
    return frac * max(float(np.abs(np.asarray(g)).max()) for g in grads.values())


def t64(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def run_oracle_step(cfg, params_np, trainable, bn_in_loss, s, x_np, eps_np, steps, training=True):
    """Oracle loss + grads in float64.  Returns (losses dict of numpy, grads dict, Ctx)."""
    params = O.to_torch(params_np, trainable)
    data = t64(x_np)
    eps = [t64(e) for e in eps_np]
    record = {}
    out, c = O.train_step_loss(cfg, s, params, bn_in_loss, data, eps, steps, training=training, record=record)
    grads = O.grads_wrt_trainables(out["loss"], c, trainable)
    losses = {k: out[k].detach().numpy() for k in ("loss", "reconstruction_loss", "kl_loss", "bn_loss", "kl_all",
                                                   "logits")}
    return losses, {k: v.detach().numpy() for k, v in grads.items()}, c, record


def compare_grads(got: dict, want: dict, tol: float, zero_frac: float = 1e-9, noise_frac: float = 1e-5, report: str = ""):
    """Per-tensor max-rel-err of every gradient; returns (worst_name, worst_err).

    A bias that feeds a training-mode BatchNorm (conv1/bias -> batch_norm2, depth_conv/bias ->
    batch_norm3, conv2/bias -> batch_norm4, ...) has an analytically ZERO gradient: BN removes any
    constant shift.  The oracle (float64) yields ~1e-17 there and any fp32 implementation yields the
    round-off of a large cancelling sum, so a *relative* error is meaningless for those tensors; they
    are instead required to stay below `noise_frac` of the largest gradient entry in the model."""
    gmax = max(float(np.abs(np.asarray(g)).max()) for g in want.values())
    worst = ("", 0.0)
    zeros, floored = [], []
    for n, w in want.items():
        w = np.asarray(w, dtype=np.float64)
        g = np.asarray(got[n], dtype=np.float64)
        wmax = float(np.abs(w).max())
        if wmax <= zero_frac * gmax:
            err = tol * float(np.abs(g).max()) / (noise_frac * gmax)  # == tol exactly at the noise bound
            zeros.append(n)
        else:
            err = max_rel_err(g, w, 1e-4 * gmax)
            if wmax < 1e-4 * gmax:  # the denominator floor (1e-4 of the largest gradient entry) is in effect: looser than
                floored.append((n, 1e-4 * gmax / wmax, max_rel_err(g, w)))  # the plain per-tensor max|d| / max|ref|
        if err > worst[1]:
            worst = (n, err)
    if report:
        # the north star's plain criterion is per-tensor max|d| / max|ref| <= tol; say exactly where this check is looser
        print(f"[{report}] {len(want)} gradient tensors: {len(zeros)} analytically zero (only bounded by {noise_frac:g} of the "
              f"largest gradient entry), {len(floored)} compared with the denominator floor in effect:")
        for n, factor, plain in sorted(floored, key=lambda t: -t[1]):
            print(f"    {n}: floor loosens x{factor:.1f}; plain max|d|/max|ref| = {plain:.2e}")
    return worst


def analytic_zero_grads(want: dict, zero_frac: float = 1e-9):
    gmax = max(float(np.abs(np.asarray(g)).max()) for g in want.values())
    return {n for n, w in want.items() if np.abs(np.asarray(w)).max() <= zero_frac * gmax}
