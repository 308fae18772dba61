"""Independent cross-checks of the oracle's restatements of third-party TensorFlow semantics that
need no TensorFlow: each one pits `oracle/nvae_oracle.py` against a second, differently-written implementation
(torch.nn.functional / torch.optim / brute-force numpy loops / hand-computed tables).

These do NOT pin the oracle against the real reference (TensorFlow 2.3 + TF-Addons + TF-Probability are not installable
in this image: parity remains *unpinned*, see README / DESIGN 5) -- they rule out transcription errors in the
restatement itself."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nvae_oracle as O


def test_batch_norm_against_torch_functional_running_statistics():
    """Keras BatchNormalization(momentum=0.05) == torch batch_norm(momentum=0.95): torch's momentum is the UPDATE weight,
    Keras' the RETAIN weight; both feed the UNBIASED batch variance into the running variance and normalise the batch
    with the BIASED one."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(6, 5, 7, 12, dtype=torch.float64, generator=g) * 3 + 1.5
    p = {"bn/gamma": torch.rand(12, dtype=torch.float64, generator=g) + 0.5,
         "bn/beta": torch.randn(12, dtype=torch.float64, generator=g),
         "bn/moving_mean": torch.randn(12, dtype=torch.float64, generator=g),
         "bn/moving_variance": torch.rand(12, dtype=torch.float64, generator=g) + 0.5}
    c = O.Ctx(p, training=True)
    y = O.batch_norm(c, "bn", x)
    rm, rv = p["bn/moving_mean"].clone(), p["bn/moving_variance"].clone()
    yt = F.batch_norm(x.permute(0, 3, 1, 2), rm, rv, p["bn/gamma"], p["bn/beta"], training=True,
                      momentum=1.0 - O.BN_MOMENTUM, eps=O.BN_EPS).permute(0, 2, 3, 1)
    assert torch.allclose(y, yt, rtol=1e-12, atol=1e-12)
    assert torch.allclose(c.new_stats["bn/moving_mean"], rm, rtol=1e-12, atol=1e-14)
    assert torch.allclose(c.new_stats["bn/moving_variance"], rv, rtol=1e-12, atol=1e-14)
    # inference: moving statistics, nothing updated
    ci = O.Ctx(p, training=False)
    yi = O.batch_norm(ci, "bn", x)
    yti = F.batch_norm(x.permute(0, 3, 1, 2), p["bn/moving_mean"], p["bn/moving_variance"], p["bn/gamma"], p["bn/beta"],
                       training=False, eps=O.BN_EPS).permute(0, 2, 3, 1)
    assert torch.allclose(yi, yti, rtol=1e-12, atol=1e-12) and not ci.new_stats


def test_adamax_against_hand_computed_table_and_keras_formula():
    """tf.keras.optimizers.Adamax dense update (Keras source `_resource_apply_dense` -> ApplyAdaMax):
        m_t = b1 m + (1-b1) g;   v_t = max(b2 v, |g|);   theta -= lr / (1 - b1^t) * m_t / (v_t + eps)
    Three steps of one scalar, worked by hand with b1=0.9, b2=0.999, eps=1e-7, lr=1e-3, theta0=1, g = (0.5, -0.25, 0.1)."""
    #   t=1: m=0.05   v=0.5     step=1e-3/0.1  *0.05  /(0.5+1e-7)
    #   t=2: m=0.02   v=0.4995  step=1e-3/0.19 *0.02  /(0.4995+1e-7)
    #   t=3: m=0.028  v=0.4990005 step=1e-3/0.271*0.028/(0.4990005+1e-7)
    table = [(0.05, 0.5, 1e-3 / 0.1 * 0.05 / (0.5 + 1e-7)),
             (0.02, 0.4995, 1e-3 / 0.19 * 0.02 / (0.4995 + 1e-7)),
             (0.028, 0.4990005, 1e-3 / 0.271 * 0.028 / (0.4990005 + 1e-7))]
    p, m, v = (torch.tensor([x], dtype=torch.float64) for x in (1.0, 0.0, 0.0))
    theta = 1.0
    for t, (g, (m_w, v_w, step)) in enumerate(zip((0.5, -0.25, 0.1), table), start=1):
        p, m, v = O.adamax_update(p, torch.tensor([g], dtype=torch.float64), m, v, t, 1e-3)
        theta -= step
        assert abs(float(m) - m_w) < 1e-15 and abs(float(v) - v_w) < 1e-15
        assert abs(float(p) - theta) < 1e-15
    # and against torch.optim.Adamax on a vector (torch puts eps inside the max -- max(b2 v, |g| + eps) -- a 1e-7
    # relative difference in the denominator, so agreement is to ~1e-6 relative, not bitwise)
    g = torch.Generator().manual_seed(1)
    w = torch.randn(50, dtype=torch.float64, generator=g).requires_grad_(True)
    opt = torch.optim.Adamax([w], lr=1e-3, betas=(0.9, 0.999), eps=1e-7)
    p, m, v = w.detach().clone(), torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64)
    for t in range(1, 6):
        grad = torch.randn(50, dtype=torch.float64, generator=g)
        w.grad = grad.clone()
        opt.step()
        p, m, v = O.adamax_update(p, grad, m, v, t, 1e-3)
    assert torch.allclose(p, w.detach(), rtol=0, atol=1e-8)


def test_cosine_decay_schedule_table():
    """tf.keras.experimental.CosineDecay(1e-3, T): lr(t) = 1e-3 * 0.5 (1 + cos(pi min(t,T)/T))."""
    T = 100
    assert O.cosine_decay_lr(0, T) == pytest.approx(1e-3)
    assert O.cosine_decay_lr(50, T) == pytest.approx(0.5e-3)
    assert O.cosine_decay_lr(25, T) == pytest.approx(1e-3 * 0.5 * (1 + np.cos(np.pi / 4)))
    assert O.cosine_decay_lr(100, T) == pytest.approx(0.0, abs=1e-18)
    assert O.cosine_decay_lr(250, T) == pytest.approx(0.0, abs=1e-18)  # clamped past decay_steps


def _brute_conv_same(x, w, stride):
    """TF SAME convolution by definition, pure loops: out = ceil(in/stride), pad_total = max((out-1) s + k - in, 0),
    pad_before = pad_total // 2 (the odd element goes AFTER), NHWC x HWIO."""
    N, H, W, Ci = x.shape
    R, S, _, Co = w.shape
    Ho, Wo = -(-H // stride), -(-W // stride)
    pt = max((Ho - 1) * stride + R - H, 0) // 2
    pl = max((Wo - 1) * stride + S - W, 0) // 2
    y = np.zeros((N, Ho, Wo, Co))
    for ho in range(Ho):
        for wo in range(Wo):
            for r in range(R):
                for s in range(S):
                    h, ww = ho * stride + r - pt, wo * stride + s - pl
                    if 0 <= h < H and 0 <= ww < W:
                        y[:, ho, wo, :] += x[:, h, ww, :] @ w[r, s]
    return y


@pytest.mark.parametrize("H,W,k,stride", [(8, 8, 3, 1), (8, 8, 3, 2), (7, 5, 3, 2), (6, 6, 5, 1), (8, 6, 1, 2),
                                          (31, 31, 1, 2), (4, 4, 5, 1)])
def test_same_padding_convolution_against_brute_force_loops(H, W, k, stride):
    rng = np.random.default_rng(H * 100 + W * 10 + k + stride)
    x = rng.normal(size=(2, H, W, 3))
    w = rng.normal(size=(k, k, 3, 4))
    y = O.conv2d(torch.as_tensor(x), torch.as_tensor(w), None, stride).numpy()
    np.testing.assert_allclose(y, _brute_conv_same(x, w, stride), rtol=1e-12, atol=1e-12)


def test_depthwise_convolution_against_brute_force_loops():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 6, 5, 4))
    w = rng.normal(size=(5, 5, 4, 1))
    b = rng.normal(size=(4,))
    y = O.depthwise_conv2d(torch.as_tensor(x), torch.as_tensor(w), torch.as_tensor(b)).numpy()
    ref = np.zeros_like(x)
    for h in range(6):
        for ww in range(5):
            for r in range(5):
                for s in range(5):
                    hh, w2 = h + r - 2, ww + s - 2
                    if 0 <= hh < 6 and 0 <= w2 < 5:
                        ref[:, h, ww, :] += x[:, hh, w2, :] * w[r, s, :, 0]
    np.testing.assert_allclose(y, ref + b, rtol=1e-12, atol=1e-12)


def test_spectral_norm_against_torch_power_iteration():
    """tfa.SpectralNormalization(power_iterations=1): v = l2n(u W^T); u' = l2n(v W); sigma = v W u'^T -- the same update
    torch.nn.utils.spectral_norm performs on the matrix W^T (its `u` is our `v` side), checked through sigma and W/sigma."""
    g = torch.Generator().manual_seed(2)
    w = torch.randn(3, 3, 5, 7, dtype=torch.float64, generator=g)
    u = torch.randn(1, 7, dtype=torch.float64, generator=g)
    c = O.Ctx({"c/kernel": w, "c/u": u}, training=True)
    wn = O.sn_kernel(c, "c")
    wm = w.reshape(-1, 7)
    v = F.normalize(u @ wm.t(), dim=1, eps=0)
    u2 = F.normalize(v @ wm, dim=1, eps=0)
    sigma = (v @ wm @ u2.t()).item()
    assert torch.allclose(wn, w / sigma, rtol=1e-13, atol=0)
    assert torch.allclose(c.new_stats["c/u"], u2, rtol=1e-13, atol=0)
    # one iteration from a random u under-estimates the top singular value; it never exceeds it
    assert 0 < sigma <= torch.linalg.matrix_norm(wm, 2).item() * (1 + 1e-12)


def test_nearest_upsample_and_activation_gradients_against_autograd_of_torch_builtins():
    x = torch.randn(2, 3, 4, 5, dtype=torch.float64)
    up = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(O.upsample_nearest2(x), up)
    assert torch.allclose(O.swish(x), F.silu(x), rtol=1e-14, atol=1e-15)
    assert torch.allclose(O.elu(x), F.elu(x), rtol=1e-14, atol=1e-15)
    assert torch.allclose(O.softclamp5(x * 10), 5.0 * torch.tanh(x * 2), rtol=1e-14, atol=1e-15)
