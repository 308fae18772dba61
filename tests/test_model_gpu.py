"""GPU parity of the residual cells and of the whole NVAE step (forward, losses, every gradient,
optimizer update, moving statistics) against the float64 oracle and the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

import helpers as H
from oracle import nvae_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL_ACT = 1e-3   # BASELINE north star: per-tensor max relative error on activations and gradients
TOL_LOSS = 1e-3  # ELBO / KL terms within 0.1 %
FP32_TOL = 1e-4  # what the fp32 CUDA-core arithmetic mode actually achieves (checked too)


def npy(t):
    return t.detach().cpu().numpy().astype(np.float64)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def _load_params(rt, rng, jitter=0.3):
    p = {}
    for v in rt.variables.values():
        if "variance" in v.name:
            a = rng.uniform(0.5, 1.5, v.shape)
        elif v.name.endswith("gamma"):
            a = rng.normal(1, 0.2, v.shape)
        elif v.name.endswith("kernel"):
            a = rng.normal(0, 1.0 / np.sqrt(max(np.prod(v.shape[:-1]), 1)), v.shape)
        else:
            a = rng.normal(0, jitter, v.shape)
        a = f32(a)
        v.assign(a)
        p[v.name] = H.t64(a).requires_grad_(v.trainable)
    return p


@pytest.mark.parametrize("kind,shape,training", [("enc", (4, 8, 8, 32), True), ("enc", (3, 7, 7, 16), False),
                                                 ("dec", (4, 4, 4, 32), True), ("dec", (2, 8, 8, 16), False),
                                                 ("enc", (8, 14, 14, 64), True), ("dec", (8, 7, 7, 128), True)])
def test_residual_cells(lib_built, kind, shape, training):
    """EncodingResidualCell (encoder.py:86-107) / GenerativeResidualCell (decoder.py:120-147) fwd + bwd."""
    from nvae_tf_b200.decoder import GenerativeResidualCell
    from nvae_tf_b200.encoder import EncodingResidualCell
    from nvae_tf_b200.runtime import DeviceTensor, Runtime
    rng = np.random.default_rng(11)
    rt = Runtime(seed=3)
    with rt:
        cell = (EncodingResidualCell if kind == "enc" else GenerativeResidualCell)(shape[-1], name="cell")
        rt.finalize()
        p = _load_params(rt, rng)
        x, dy = f32(rng.normal(0, 1, shape)), f32(rng.normal(0, 1, shape))
        xt = DeviceTensor(torch.as_tensor(x.astype(np.float32)).to(rt.device))
        with rt.gradient_tape() as tape:
            y = cell(xt, training=training)
        y.grad = torch.as_tensor(dy.astype(np.float32)).to(rt.device)
        rt.backward(tape)
    c = O.Ctx(p, training)
    xo = H.t64(x).requires_grad_(True)
    yo = (O.encoding_residual_cell if kind == "enc" else O.generative_residual_cell)(c, "cell", xo)
    leaves = {n: (c.new_stats.get(n, t) if n.endswith("/kernel") else t) for n, t in p.items() if t.requires_grad}
    grads = torch.autograd.grad(yo, [xo] + list(leaves.values()), H.t64(dy), allow_unused=True)
    assert H.max_rel_err(npy(y.data), yo.detach().numpy()) < FP32_TOL
    assert H.max_rel_err(npy(xt.grad), grads[0].numpy()) < FP32_TOL
    want = {n: (g.numpy() if g is not None else np.zeros(rt.variables[n].shape))
            for (n, _), g in zip(leaves.items(), grads[1:])}
    worst = H.compare_grads({n: npy(rt.variables[n].grad) for n in want}, want, FP32_TOL)
    assert worst[1] < FP32_TOL, worst
    if training:  # SN normalised the kernels in place and advanced u; BN moving statistics moved
        for n, t in c.new_stats.items():
            v = rt.variables[n]
            assert H.max_rel_err(npy(v.value), t.detach().numpy()) < FP32_TOL, n


def _make_model(cfg, batch, training, **kw):
    from nvae_tf_b200.models import NVAE, Adamax, CosineDecay
    m = NVAE(**H.mirror_kwargs(cfg, batch), training=training, **kw)
    m.compile(optimizer=Adamax(learning_rate=CosineDecay(1e-3, 1000)), run_eagerly=True)
    return m


def _compare_step(m, f_losses, f_acts, f_grads, f_new, out, recon_t, tol):
    rt = m.rt
    total = float(out["loss"].item())
    assert abs(total - float(f_losses["loss"])) <= TOL_LOSS * abs(float(f_losses["loss"]))
    assert H.max_rel_err(npy(out["reconstruction_loss"]), f_losses["reconstruction_loss"]) < TOL_LOSS
    assert H.max_rel_err(npy(out["kl_loss"]), f_losses["kl_loss"], floor=1e-3) < TOL_LOSS
    assert abs(float(out["bn_loss"].item()) - float(f_losses["bn_loss"])) <= TOL_LOSS * float(f_losses["bn_loss"])
    assert H.max_rel_err(npy(m.decoder.sampler.kl_all), f_losses["kl_all"]) < TOL_LOSS
    assert H.max_rel_err(npy(recon_t), f_losses["logits"]) < tol
    worst = H.compare_grads(rt.named_grads(), f_grads, tol)
    assert worst[1] < tol, worst
    for n, v in f_new.items():
        assert H.max_rel_err(npy(rt.variables[n].value), v) < tol, n


@pytest.mark.parametrize("name", ["tiny_train_balanced", "tiny_infer_beta1"])
def test_full_step_against_golden_fixture(lib_built, name):
    f = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = H.oracle_cfg()
    training, steps = bool(f["training"]), int(f["steps"])
    x = f["x"]
    m = _make_model(cfg, x.shape[0], training)
    rt = m.rt
    rt.load_named({k[len("param/"):]: f[k] for k in f.files if k.startswith("param/")})
    n_groups = m.decoder.sampler.n_groups
    rt.inject_eps([f[f"eps/{i}"] for i in range(n_groups)])
    m.steps = steps
    captured = {}
    orig_call = m.postprocess.__class__.__call__

    def spy(self, inputs, training=False):
        out = orig_call(self, inputs, training)
        captured["logits"] = out
        return out
    m.postprocess.__class__.__call__ = spy
    try:
        out = m.train_step(x, apply_gradients=False)
    finally:
        m.postprocess.__class__.__call__ = orig_call
    losses = {k[5:]: f[k] for k in f.files if k.startswith("loss/")}
    grads = {k[5:]: f[k] for k in f.files if k.startswith("grad/")}
    new = {k[4:]: f[k] for k in f.files if k.startswith("new/")}
    _compare_step(m, losses, None, grads, new, out, captured["logits"].data, FP32_TOL)
    assert m.steps == steps + 1


@pytest.mark.parametrize("batch,steps,training,f16x3", [(6, 20000, True, False), (5, 80000, False, False),
                                                        (6, 20000, True, True)])
def test_default_config_step_against_live_oracle(lib_built, batch, steps, training, f16x3, monkeypatch):
    """train.py defaults (40.1 M parameters, 15 latent groups), small batch so the float64 oracle runs in seconds.
    f16x3: every eligible convolution is forced onto the 3xFP16 path the batch-144 step uses for its six large GEMMs
    (at this batch none would reach the 20 GFLOP threshold), so the arithmetic is checked through the whole model."""
    monkeypatch.setenv("NVAE_F16X3_MIN_GFLOP", "0" if f16x3 else "1e9")
    cfg = O.NVAEConfig()
    params, trainable, bnl, s = O.build_params(cfg, seed=1, jitter=0.05)
    params = {k: f32(v) for k, v in params.items()}
    x = O.make_images(cfg, batch, seed=1).numpy()
    eps = [f32(e.numpy()) for e in O.make_eps(s, batch, seed=1)]
    losses, grads, c, record = H.run_oracle_step(cfg, params, trainable, bnl, s, x, eps, steps, training)
    m = _make_model(cfg, batch, training)
    m.rt.load_named(params)
    m.rt.inject_eps(eps)
    m.steps = steps
    out = m.train_step(x, apply_gradients=False)
    assert abs(float(out["loss"].item()) - float(losses["loss"])) <= TOL_LOSS * abs(float(losses["loss"]))
    assert H.max_rel_err(npy(out["reconstruction_loss"]), losses["reconstruction_loss"]) < TOL_LOSS
    assert H.max_rel_err(npy(out["kl_loss"]), losses["kl_loss"], floor=1e-3) < TOL_LOSS
    assert H.max_rel_err(npy(m.decoder.sampler.kl_all), losses["kl_all"]) < TOL_LOSS
    worst = H.compare_grads(m.rt.named_grads(), grads, TOL_ACT, report=f"default config, batch {batch}, f16x3={f16x3}")
    assert worst[1] < TOL_ACT, worst
    new = {k: v.detach().numpy() for k, v in c.new_stats.items()}
    for n, v in new.items():
        assert H.max_rel_err(npy(m.rt.variables[n].value), v) < TOL_ACT, n


def test_cifar_shape_three_scales_against_live_oracle(lib_built):
    """BASELINE configs[4]: 32x32x3 input, deeper hierarchy (three latent scales: 8x8, 4x4, 2x2).  The head keeps the
    reference's single output channel (postprocess.py:29), so the logits broadcast over the 3 input channels in
    the Bernoulli log-likelihood (README.md:25-27) -- parity with that behaviour, not a redesign."""
    cfg = O.NVAEConfig(n_encoder_channels=8, n_decoder_channels=8, n_latent_per_group=8, n_groups_per_scale=(2, 3, 4),
                       n_preprocess_cells=2, n_post_process_cells=2, image_channels=3, n_total_iterations=100)
    batch, steps = 4, 20
    params, trainable, bnl, s = O.build_params(cfg, seed=5, jitter=0.1)
    params = {k: f32(v) for k, v in params.items()}
    x = O.make_images(cfg, batch, seed=5).numpy()
    assert x.shape == (batch, 32, 32, 3)
    eps = [f32(e.numpy()) for e in O.make_eps(s, batch, seed=5)]
    losses, grads, c, record = H.run_oracle_step(cfg, params, trainable, bnl, s, x, eps, steps, True)
    m = _make_model(cfg, batch, True)
    assert m.decoder.sampler.n_groups == 9
    m.rt.load_named(params)
    m.rt.inject_eps(eps)
    m.steps = steps
    out = m.train_step(x, apply_gradients=False)
    assert abs(float(out["loss"].item()) - float(losses["loss"])) <= TOL_LOSS * abs(float(losses["loss"]))
    assert H.max_rel_err(npy(out["reconstruction_loss"]), losses["reconstruction_loss"]) < TOL_LOSS
    assert H.max_rel_err(npy(m.decoder.sampler.kl_all), losses["kl_all"]) < TOL_LOSS
    worst = H.compare_grads(m.rt.named_grads(), grads, TOL_ACT)
    assert worst[1] < TOL_ACT, worst


def test_optimizer_update_and_second_step(lib_built):
    """apply_gradients: Adamax + CosineDecay on the flat arena == per-variable oracle updates; step 2 still agrees."""
    cfg = H.oracle_cfg()
    params, trainable, bnl, s = O.build_params(cfg, seed=6, jitter=0.1)
    params = {k: f32(v) for k, v in params.items()}
    x = O.make_images(cfg, 4, seed=6).numpy()
    m = _make_model(cfg, 4, True)
    m.rt.load_named(params)
    cur = dict(params)
    mom = {n: (np.zeros_like(params[n]), np.zeros_like(params[n])) for n in trainable}
    noisy = set()
    for step in range(2):
        eps = [f32(e.numpy()) for e in O.make_eps(s, 4, seed=10 + step)]
        m.rt.inject_eps(eps)
        m.steps = 5 + step
        out = m.train_step(x)
        losses, grads, c, _ = H.run_oracle_step(cfg, cur, trainable, bnl, s, x, eps, 5 + step, True)
        assert abs(float(out["loss"].item()) - float(losses["loss"])) <= TOL_LOSS * abs(float(losses["loss"]))
        for n, v in c.new_stats.items():
            cur[n] = v.detach().numpy()
        noisy |= H.analytic_zero_grads(grads)
        lr = O.cosine_decay_lr(step, 1000)
        for n in trainable:
            p, mm, vv = O.adamax_update(H.t64(cur[n]), H.t64(grads[n]), H.t64(mom[n][0]), H.t64(mom[n][1]), step + 1, lr)
            cur[n], mom[n] = p.numpy(), (mm.numpy(), vv.numpy())
    got = m.rt.named_values()
    # Adamax's first updates are ~lr*sign(g) per element, so an element whose gradient is round-off noise
    # (|g| ~ 1e-7 of the tensor's scale; whole tensors for a bias in front of a training BN, whose gradient is
    # analytically zero) moves by ~lr in a noise-determined direction -- in TensorFlow as here.  Hence:
    # every element stays within the total step size, and all but a sliver agree tightly.
    lr_total = sum(O.cosine_decay_lr(i, 1000) / (1 - 0.9 ** (i + 1)) for i in range(2))
    loose = noisy | {n for n in cur if n.endswith("moving_mean")}
    assert noisy
    for n in cur:
        diff = np.abs(got[n] - cur[n])
        assert diff.max() <= 2.0 * lr_total + 1e-6, (n, diff.max())
        if n not in loose:
            bad = float((diff > 1e-5 * np.abs(cur[n]).max() + 1e-7).mean())
            assert bad < 5e-3, (n, bad)


def test_nll_path_and_public_loss_methods(lib_built):
    """model(x, nll=True) log q / log p sums (decoder.py:69-102), calculate_kl_loss both branches, crop."""
    cfg = H.oracle_cfg()
    params, trainable, bnl, s = O.build_params(cfg, seed=8, jitter=0.1)
    params = {k: f32(v) for k, v in params.items()}
    x = O.make_images(cfg, 3, seed=8).numpy()
    eps = [f32(e.numpy()) for e in O.make_eps(s, 3, seed=8)]
    m = _make_model(cfg, 3, False)
    m.rt.load_named(params)
    m.rt.inject_eps(eps)
    recon, z_params, log_p, log_q = m(x, nll=True)
    c = O.Ctx(O.to_torch(params, []), False, [H.t64(e) for e in eps])
    lo, zo, lpo, lqo = O.nvae_call(c, s, H.t64(x), nll=True)
    assert H.max_rel_err(npy(recon.data), lo.numpy()) < FP32_TOL
    assert H.max_rel_err(npy(log_p), lpo.numpy()) < TOL_LOSS and H.max_rel_err(npy(log_q), lqo.numpy()) < TOL_LOSS
    for balancing in (True, False):
        kl = m.calculate_kl_loss(z_params, balancing)
        assert H.max_rel_err(npy(kl), O.calculate_kl_loss(cfg, zo, balancing)[0].numpy()) < TOL_LOSS
    for crop in (False, True):
        r = m.calculate_recon_loss(x, recon, crop_output=crop)
        assert H.max_rel_err(npy(r), O.calculate_recon_loss(H.t64(x), lo, crop).numpy()) < TOL_LOSS
    assert H.max_rel_err(npy(z_params[2].enc_sigma), zo[2].enc_sigma.numpy()) < FP32_TOL
    assert len(z_params) == 4 and z_params[0].dec_sigma.min().item() == 1.0


def test_sampling_matches_oracle(lib_built):
    """NVAE.sample (models.py:137-178): temperature only scales z0, inference BN, no SN."""
    cfg = H.oracle_cfg()
    params, trainable, bnl, s = O.build_params(cfg, seed=9, jitter=0.1)
    params = {k: f32(v) for k, v in params.items()}
    n = 5
    eps = [f32(e.numpy()) for e in O.make_eps(s, n, seed=9)]
    m = _make_model(cfg, n, False)
    m.rt.load_named(params)
    extra = [eps[-1] * 0.5, eps[-1] * -0.25]  # the two PPL draws z1, z2 (models.py:175-176)
    m.rt.inject_eps(eps + extra)
    before = m.rt.named_values()
    images, last_s, z1, z2 = m.sample(n_samples=n, temperature=0.7)
    probs, _ = O.sample(cfg, s, O.to_torch(params, []), n, 0.7, [H.t64(e) for e in eps])
    assert H.max_rel_err(npy(images), probs.numpy()) < FP32_TOL
    assert images.shape == (n, 32, 32, 1) and z1.shape == z2.shape == (n, 8, 8, cfg.n_latent_per_group)
    after = m.rt.named_values()
    assert all(np.array_equal(before[k], after[k]) for k in before)  # sampling never mutates weights or statistics
    img2 = m.sample_with_z(z1, last_s)
    assert img2.shape == images.shape and float(img2.min()) >= 0 and float(img2.max()) <= 1


def test_cuda_graph_replay_equals_eager(lib_built):
    """The captured whole-step graph IS the eager step: after 5 steps every parameter, every Adamax slot and every
    moving statistic is bit-identical (Philox epsilons are keyed by the device iteration counter + the group index, and
    capture leaves the model untouched)."""
    cfg = H.oracle_cfg()
    x = torch.as_tensor(O.make_images(cfg, 4, seed=3).numpy().astype(np.float32))
    res = []
    for mode in ("eager", "graph"):
        m = _make_model(cfg, 4, True, seed=5)
        m.steps = 7
        if mode == "eager":
            outs = [m.train_step(x.cuda()) for _ in range(5)]
            loss = outs[-1]["loss"].item()
        else:
            m._sync_counters()  # `m.steps = 7` above reaches the device counter on the next step / capture; do it now
            before = (m.rt.params.clone(), m.rt.state.clone(), m._counters.clone())
            static_in, replay = m.capture_train_step((4, 32, 32, 1), warmup=2)
            torch.cuda.synchronize()
            assert torch.equal(before[0], m.rt.params) and torch.equal(before[1], m.rt.state)  # capture mutates nothing
            assert torch.equal(before[2], m._counters) and m.steps == 7
            assert float(m._m.abs().max()) == 0.0 and float(m._v.abs().max()) == 0.0
            static_in.copy_(x)
            for _ in range(5):
                out = replay()
            loss = out["loss"].item()
            assert m.graph_launches > 100
        torch.cuda.synchronize()
        assert m.steps == 12
        res.append((loss, m.rt.params.clone(), m.rt.state.clone(), m._m.clone(), m._v.clone()))
    assert res[0][0] == res[1][0]
    for a, b in zip(res[0][1:], res[1][1:]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["1", "2"])
def test_bn_apply_inside_the_1x1_convolutions_is_the_same_step(lib_built, mode, monkeypatch):
    """NVAE_FUSE_BN_CONV (opt-in): the BN-apply (1) and BN + swish (2) in front of the cells' 1x1 convolutions run in the
    convolution's operand path.  Same values into the same tile arithmetic: three optimizer steps end bit-identical to
    the default step."""
    cfg = H.oracle_cfg()
    x = torch.as_tensor(O.make_images(cfg, 4, seed=3).numpy().astype(np.float32))
    res = []
    for fuse in ("0", mode):
        monkeypatch.setenv("NVAE_FUSE_BN_CONV", fuse)
        m = _make_model(cfg, 4, True, seed=5)
        for _ in range(3):
            out = m.train_step(x.cuda())
        torch.cuda.synchronize()
        res.append((out["loss"].item(), m.rt.params.clone(), m.rt.state.clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


def test_graph_replay_follows_epoch_based_warmup(lib_built):
    """train.py default: --step_based_warmup off, beta = min(epoch / (0.3 n_total), 1) moves only in on_epoch_begin
    (models.py:121-122).  The replayed graph must see the new epoch (it used to keep the capture-time beta)."""
    cfg = H.oracle_cfg(step_based_warmup=False, n_total_iterations=20)
    x = torch.as_tensor(O.make_images(cfg, 4, seed=4).numpy().astype(np.float32)).cuda()
    runs = []
    for mode in ("eager", "graph"):
        m = _make_model(cfg, 4, True, seed=6)
        step = (lambda: m.train_step(x)) if mode == "eager" else None
        if mode == "graph":
            static_in, replay = m.capture_train_step((4, 32, 32, 1))
            static_in.copy_(x)
            step = replay
        seen = []
        for epoch in (0, 0, 3, 3, 9):
            m.on_epoch_begin(epoch)
            out = step()
            torch.cuda.synchronize()
            seen.append((float(m._hyper[0].item()), out["kl_loss"].clone(), float(out["loss"].item())))
        runs.append(seen)
    betas = [b for b, _, _ in runs[1]]
    assert betas == [0.0, 0.0, 0.5, 0.5, 1.0], betas
    for (b0, k0, l0), (b1, k1, l1) in zip(*runs):
        assert b0 == b1 and torch.equal(k0, k1) and l0 == l1
    assert float(runs[1][2][1].abs().max()) > 0.0  # KL really is weighted in once beta > 0


def test_checkpoint_round_trip_resumes_the_schedules(lib_built, tmp_path):
    """save_weights / load_weights carry the variables AND the optimizer slots, iteration count and warm-up position
    (Keras TF-format checkpoints do; train.py:133-135): a resumed model takes bit-identical next steps."""
    cfg = H.oracle_cfg()
    x = torch.as_tensor(O.make_images(cfg, 4, seed=5).numpy().astype(np.float32)).cuda()
    a = _make_model(cfg, 4, True, seed=8)
    for _ in range(3):
        a.train_step(x)
    path = str(tmp_path / "epoch_3")
    a.save_weights(path)
    b = _make_model(cfg, 4, True, seed=99)  # different init: everything must come from the file
    b.load_weights(path)
    assert b.steps == a.steps == 3 and int(b._counters[1].item()) == 3
    assert torch.equal(a.rt.params, b.rt.params) and torch.equal(a.rt.state, b.rt.state)
    assert torch.equal(a._m, b._m) and torch.equal(a._v, b._v)
    b.rt.philox_seed = a.rt.philox_seed
    oa, ob = a.train_step(x), b.train_step(x)
    torch.cuda.synchronize()
    assert oa["loss"].item() == ob["loss"].item() and torch.equal(a.rt.params, b.rt.params)
    assert float(b._hyper[3].item()) == 4.0  # Adamax bias-correction step t continued at 4, not restarted at 1


def test_train_function_handles_the_short_last_batch(lib_built):
    """The reference's last batch of an epoch is smaller (96 of 144): make_train_function captures a second
    graph for the new shape and the result equals the eager step on the same state."""
    cfg = H.oracle_cfg()
    x6 = O.make_images(cfg, 6, seed=6).numpy().astype(np.float32)
    x4 = O.make_images(cfg, 4, seed=7).numpy().astype(np.float32)
    m = _make_model(cfg, 6, True, seed=9)
    e = _make_model(cfg, 6, True, seed=9)
    fn = m.make_train_function((6, 32, 32, 1))
    r1, r2, r3 = fn(x6), fn(x4), fn(x6)
    o = [e.train_step(torch.as_tensor(v).cuda()) for v in (x6, x4, x6)]
    torch.cuda.synchronize()
    assert len(fn.captured_shapes) == 2 and r2["kl_loss"].shape == (4,)
    assert r1["loss"] == o[0]["loss"].item() and r2["loss"] == o[1]["loss"].item() and r3["loss"] == o[2]["loss"].item()
    assert torch.equal(m.rt.params, e.rt.params)


def test_neg_log_likelihood_matches_oracle(lib_built):
    """evaluate.py:111-123: k-attempt importance-weighted bound with the 28x28 crop, logsumexp on the device."""
    from nvae_tf_b200.evaluate import batch_neg_log_likelihood, neg_log_likelihood
    cfg = H.oracle_cfg()
    params, trainable, bnl, s = O.build_params(cfg, seed=12, jitter=0.1)
    params = {k: f32(v) for k, v in params.items()}
    x = O.make_images(cfg, 5, seed=12).numpy()
    K = 3
    eps = [[f32(e.numpy()) for e in O.make_eps(s, 5, seed=20 + k)] for k in range(K)]
    m = _make_model(cfg, 5, False)
    m.rt.load_named(params)
    m.rt.inject_eps([e for att in eps for e in att])  # attempt k consumes the k-th run of per-group epsilons
    got = float(batch_neg_log_likelihood(m, x, n_attempts=K).item())
    logs = []
    for k in range(K):
        c = O.Ctx(O.to_torch(params, []), False, [H.t64(e) for e in eps[k]])
        lo, _, lp, lq = O.nvae_call(c, s, H.t64(x), nll=True)
        logs.append(-O.calculate_recon_loss(H.t64(x), lo, True) - lq + lp)
    want = float(-(torch.logsumexp(torch.stack(logs), 0) - np.log(K)).mean())
    assert abs(got - want) <= TOL_LOSS * abs(want), (got, want)
    m.rt.inject_eps(None)
    metric = neg_log_likelihood(m, [(x, None), (x[:3], None)], n_attempts=2)
    assert np.isfinite(metric.mean) and metric.stddev >= 0.0


def test_sample_graph_replay(lib_built):
    """NVAE.capture_sample: the whole ancestral-sampling pass as one CUDA graph; each replay draws fresh epsilons and
    picks up retrained weights."""
    cfg = H.oracle_cfg()
    m = _make_model(cfg, 4, False, seed=4)
    replay = m.capture_sample(n_samples=8, temperature=0.7)
    a = replay()[0].clone()
    b = replay()[0].clone()
    torch.cuda.synchronize()
    assert a.shape == (8, 32, 32, 1) and float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    assert not torch.equal(a, b)  # new noise per replay
    assert m.sample_graph_kernels > 20
    # same counter value + same weights -> same images as the eager call
    m._sample_counters.zero_()
    g0 = replay()[0].clone()
    prev, m.rt.counters = m.rt.counters, torch.zeros(2, dtype=torch.int64, device=m.rt.device)
    m.rt.eps_i = 0
    e0 = m.sample(n_samples=8, temperature=0.7)[0]
    m.rt.counters = prev
    torch.cuda.synchronize()
    assert torch.equal(g0, e0)
