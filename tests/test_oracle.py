"""CPU tests of the oracle (oracle/nvae_oracle.py).  PARITY UNPINNED: the reference ships no golden
vectors and TensorFlow is unavailable, so the oracle is pinned by (1) structural facts of the
reference that can be derived by hand from its constructors, (2) finite-difference checks of its
gradients, (3) closed-form identities, (4) the committed fixtures (drift guard)."""
import math
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import nvae_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_default_config_structure():
    # SURVEY Appendix B (symbolic walk of models.py:16-87 with train.py:145-216 defaults)
    cfg = O.NVAEConfig()
    params, trainable, bnl, s = O.build_params(cfg)
    assert sum(params[k].size for k in trainable) == 40_128_893
    assert sum(1 for k in params if k.endswith("/u")) == 163            # SpectralNormalization wrappers
    assert sum(1 for k in params if k.endswith("/gamma")) == 128        # BatchNormalization layers
    assert len(bnl) == 88                                               # reached by calculate_bn_loss
    assert sum(1 for k in params if k.endswith("depthwise_kernel")) == 14
    assert sum(1 for k in params if k.endswith("/dense1/kernel")) == 41  # SqueezeExcitation blocks
    assert len(s.samplers) == 15
    assert params["decoder/h"].shape == (4, 4, 32)
    assert params["decoder/groups/0/conv/kernel"].shape == (1, 1, 52, 256)


def test_same_padding_is_tf_asymmetric():
    assert O.same_pad(32, 3, 2) == (16, 0, 1)  # extra padding goes after
    assert O.same_pad(31, 1, 2) == (16, 0, 0)
    assert O.same_pad(8, 5, 1) == (8, 2, 2)
    x = torch.arange(16.0, dtype=torch.float64).reshape(1, 4, 4, 1)
    w = torch.ones(3, 3, 1, 1, dtype=torch.float64)
    y = O.conv2d(x, w, None, stride=2)
    # windows start at rows/cols 0 and 2, the window at 2 hangs over the bottom/right edge
    assert y.shape == (1, 2, 2, 1)
    assert y[0, 0, 0, 0].item() == sum([0, 1, 2, 4, 5, 6, 8, 9, 10])
    assert y[0, 1, 1, 0].item() == sum([10, 11, 14, 15])


def test_kl_alphas_default():
    a = O.kl_alphas(O.NVAEConfig())
    assert a.shape == (15,)
    np.testing.assert_allclose(a, [1.0] * 10 + [8.0] * 5)  # models.py:227-237 with groups [5,10]


def test_kl_closed_form_and_sampler_identities():
    torch.manual_seed(0)
    a, b, c, d = (torch.randn(2, 3, 3, 4, dtype=torch.float64) for _ in range(4))
    mu_q, sig_q = O.softclamp5(a + c), torch.exp(O.softclamp5(b + d)) + 1e-2
    mu_p, sig_p = O.softclamp5(c), torch.exp(O.softclamp5(d)) + 1e-2
    kl = O.kl_per_group([O.DistributionParams(mu_q, sig_q, mu_p, sig_p)])[0]
    ref = torch.distributions.kl_divergence(torch.distributions.Normal(mu_q, sig_q),
                                            torch.distributions.Normal(mu_p, sig_p)).sum(dim=(1, 2, 3))
    torch.testing.assert_close(kl, ref)
    z = mu_q + 0.3 * sig_q
    ref_lp = torch.distributions.Normal(mu_p, sig_p).log_prob(z)
    torch.testing.assert_close(O.calculate_log_p(z, mu_p, sig_p), ref_lp)


def test_bernoulli_matches_torch_bce():
    torch.manual_seed(1)
    l = torch.randn(3, 32, 32, 1, dtype=torch.float64) * 3
    x = (torch.rand(3, 32, 32, 1, dtype=torch.float64) < 0.2).double()
    ref = torch.nn.functional.binary_cross_entropy_with_logits(l, x, reduction="none").sum(dim=(1, 2, 3))
    torch.testing.assert_close(O.calculate_recon_loss(x, l), ref)
    assert O.calculate_recon_loss(x, l, crop_output=True).shape == (3,)


def test_spectral_norm_one_power_iteration():
    rng = np.random.default_rng(0)
    w = torch.as_tensor(rng.normal(size=(3, 3, 8, 6)))
    u = torch.as_tensor(rng.normal(size=(1, 6)))
    c = O.Ctx({"c/kernel": w, "c/u": u}, training=True)
    wn = O.sn_kernel(c, "c")
    wm = w.reshape(-1, 6)
    v = u @ wm.T
    v = v / v.norm()
    u2 = v @ wm
    u2 = u2 / u2.norm()
    sigma = (v @ wm @ u2.T).item()
    torch.testing.assert_close(wn, w / sigma)
    torch.testing.assert_close(c.new_stats["c/u"], u2)
    # sigma from one iteration lower-bounds the true spectral norm
    assert sigma <= torch.linalg.matrix_norm(wm, 2).item() + 1e-9
    # inference: kernel untouched
    assert O.sn_kernel(O.Ctx({"c/kernel": w, "c/u": u}, training=False), "c") is w


def test_batch_norm_training_and_moving_stats():
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.normal(2.0, 3.0, size=(4, 5, 5, 8)))
    p = {"b/gamma": torch.full((8,), 1.5, dtype=torch.float64), "b/beta": torch.full((8,), 0.25, dtype=torch.float64),
         "b/moving_mean": torch.zeros(8, dtype=torch.float64), "b/moving_variance": torch.ones(8, dtype=torch.float64)}
    c = O.Ctx(p, training=True)
    y = O.batch_norm(c, "b", x)
    flat = y.reshape(-1, 8)
    torch.testing.assert_close(flat.mean(0), torch.full((8,), 0.25, dtype=torch.float64))
    n = 100
    mean, var = x.reshape(-1, 8).mean(0), x.reshape(-1, 8).var(0, unbiased=False)
    torch.testing.assert_close(c.new_stats["b/moving_mean"], 0.95 * mean)
    torch.testing.assert_close(c.new_stats["b/moving_variance"], 0.05 + 0.95 * var * n / (n - 1))
    yi = O.batch_norm(O.Ctx(p, training=False), "b", x)
    torch.testing.assert_close(yi, x / math.sqrt(1 + 1e-5) * 1.5 + 0.25)


def _fd_loss(cfg, s, pt, bnl, x, eps, steps, training, frozen_coeff=None):
    """Total loss with SN bypassed (kernels used as given: the straight-through leaf of SURVEY A.2) and,
    when balancing, the stop-gradient coefficients of models.py:215-217 frozen to `frozen_coeff`."""
    c = O.Ctx(pt, training, [H.t64(e) for e in eps])
    orig = O.sn_kernel
    O.sn_kernel = lambda c_, name: c_.p[name + "/kernel"]
    try:
        logits, zp, _, _ = O.nvae_call(c, s, H.t64(x))
    finally:
        O.sn_kernel = orig
    recon = O.calculate_recon_loss(H.t64(x), logits)
    beta = O.beta_schedule(cfg, steps)
    kl_all = O.kl_per_group(zp)
    coeff = None
    if beta < 1:
        alphas = torch.as_tensor(O.kl_alphas(cfg), dtype=kl_all.dtype)
        coeff = kl_all.abs().mean(dim=1) + 0.01
        coeff = coeff / alphas * coeff.sum()
        coeff = (coeff / coeff.mean()).detach()
        kl = (kl_all * (frozen_coeff if frozen_coeff is not None else coeff)[:, None]).sum(0)
    else:
        kl = kl_all.sum(0)
    total = (recon + beta * kl).mean() + O.calculate_bn_loss(c, bnl, cfg.sr_lambda)
    return total, coeff


@pytest.mark.parametrize("training,steps", [(True, 10), (True, 1000), (False, 1000)])
def test_oracle_gradients_by_finite_differences(training, steps):
    """Central differences in float64 on a handful of coordinates of every parameter kind."""
    cfg = H.oracle_cfg(n_groups_per_scale=(1, 2), n_preprocess_cells=2, n_post_process_cells=1)
    params, trainable, bnl, s = O.build_params(cfg, seed=5, jitter=0.1)
    x = O.make_images(cfg, 2, seed=5).numpy()
    eps = [e.numpy() for e in O.make_eps(s, 2, seed=5)]
    pt = O.to_torch(params, trainable)
    total, coeff = _fd_loss(cfg, s, pt, bnl, x, eps, steps, training)
    names = ["preprocess/stem/kernel", "encoder/groups/0/cells/0/conv1/kernel",
             "encoder/groups/0/cells/0/batch_norm2/beta", "decoder/groups/1/cells/0/depth_conv/depthwise_kernel",
             "decoder/groups/1/cells/0/se/dense1/kernel", "decoder/sampler/enc_sampler/1/kernel",
             "decoder/sampler/dec_sampler/1/conv/bias", "decoder/h", "postprocess/cells/0/node/cbs2/conv/kernel",
             "decoder/groups/2/conv/kernel"]
    grads = torch.autograd.grad(total, [pt[n] for n in names])
    rng = np.random.default_rng(0)
    for n, g in zip(names, grads):
        flat = params[n].reshape(-1)
        for idx in rng.choice(flat.size, size=min(3, flat.size), replace=False):
            h = 1e-5
            old = flat[idx]
            flat[idx] = old + h
            lp = _fd_loss(cfg, s, O.to_torch(params, trainable), bnl, x, eps, steps, training, coeff)[0].item()
            flat[idx] = old - h
            lm = _fd_loss(cfg, s, O.to_torch(params, trainable), bnl, x, eps, steps, training, coeff)[0].item()
            flat[idx] = old
            fd = (lp - lm) / (2 * h)
            an = g.reshape(-1)[idx].item()
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(fd)) + 1e-6, (n, idx, fd, an)


@pytest.mark.parametrize("name", ["tiny_train_balanced", "tiny_infer_beta1"])
def test_oracle_reproduces_golden(name):
    f = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = H.oracle_cfg()
    _, trainable, bnl, s = O.build_params(cfg, seed=0)
    params = {k[len("param/"):]: f[k].astype(np.float64) for k in f.files if k.startswith("param/")}
    eps = [f[f"eps/{i}"] for i in range(len(s.samplers))]
    losses, grads, c, record = H.run_oracle_step(cfg, params, trainable, bnl, s, f["x"], eps, int(f["steps"]),
                                                 bool(f["training"]))
    for k in ("loss", "reconstruction_loss", "kl_loss", "bn_loss", "kl_all"):
        np.testing.assert_allclose(losses[k], f["loss/" + k], rtol=2e-5, atol=1e-6)
    for k in (k for k in f.files if k.startswith("act/")):
        assert H.max_rel_err(record[k[4:]].detach().numpy(), f[k]) < 2e-5
    floor = H.grad_floor(grads)
    worst = max(H.max_rel_err(grads[k[5:]], f[k], floor) for k in f.files if k.startswith("grad/"))
    assert worst < 1e-4, worst


def test_data_parallel_emulation_matches_gradient_average():
    """N replicas with per-replica BN statistics, gradients averaged == what the NCCL path computes."""
    cfg = H.oracle_cfg(n_groups_per_scale=(1, 1), n_preprocess_cells=2, n_post_process_cells=1)
    params, trainable, bnl, s = O.build_params(cfg, seed=2, jitter=0.05)
    x = O.make_images(cfg, 4, seed=2).numpy()
    eps = [e.numpy() for e in O.make_eps(s, 4, seed=2)]
    gs = []
    for r in range(2):
        sl = slice(2 * r, 2 * r + 2)
        _, g, _, _ = H.run_oracle_step(cfg, params, trainable, bnl, s, x[sl], [e[sl] for e in eps], 50, True)
        gs.append(g)
    avg = {k: 0.5 * (gs[0][k] + gs[1][k]) for k in gs[0]}
    # the bn_loss term depends on the weights only: identical on both replicas, so the average keeps it whole
    k = "encoder/groups/0/cells/0/batch_norm1/gamma"
    assert np.isfinite(avg[k]).all()
    _, gfull, _, _ = H.run_oracle_step(cfg, params, trainable, bnl, s, x, eps, 50, True)
    # per-replica BN is NOT the same computation as full-batch BN: the averages must differ somewhere
    assert max(H.max_rel_err(avg[n], gfull[n]) for n in avg) > 1e-6
