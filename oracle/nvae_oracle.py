"""CPU oracle for the NVAE-TF hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED: the reference (stevensdavid/nvae-tf) ships no tests, golden vectors or
fixtures, and TensorFlow / TF-Addons / TF-Probability are not installable in the build
image, so this restatement cannot be checked against outputs of the reference itself.
It follows the reference files line by line (citations are `file:line` into the
reference tree) and restates the published semantics of the third-party ops it calls
(TensorFlow 2.3.0 `requirements.txt:41`; tensorflow_addons / tensorflow_probability are
unpinned by the reference).  Its analytic gradients (torch autograd, float64) are
self-checked by central finite differences in `tests/test_oracle.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl
reference` legs may import this module.  The product package `nvae_tf_b200` never does.

Layout conventions are TensorFlow's: activations NHWC, conv kernels HWIO, depthwise
kernels [5,5,C,1], dense kernels [in,out].  Parameters live in a flat dict keyed by the
reference's attribute paths (e.g. ``encoder/groups/0/cells/0/conv1/kernel``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_MOMENTUM = 0.05  # common.py:148, encoder.py:91, decoder.py:125 (Keras "momentum" = retain factor)
BN_EPS = 1e-5


# --------------------------------------------------------------------------------------
# configuration (constructor kwargs of NVAE, models.py:17-36; defaults train.py:145-216)
# --------------------------------------------------------------------------------------
@dataclass
class NVAEConfig:
    n_encoder_channels: int = 32
    n_decoder_channels: int = 32
    res_cells_per_group: int = 1
    n_preprocess_blocks: int = 2
    n_preprocess_cells: int = 3
    n_latent_per_group: int = 20
    n_groups_per_scale: Tuple[int, ...] = (5, 10)
    n_postprocess_blocks: int = 2
    n_post_process_cells: int = 3
    sr_lambda: float = 0.01
    scale_factor: int = 2
    total_epochs: int = 400
    n_total_iterations: int = 417 * 400
    step_based_warmup: bool = True
    image_size: int = 32
    image_channels: int = 1

    @property
    def n_latent_scales(self) -> int:
        return len(self.n_groups_per_scale)


# --------------------------------------------------------------------------------------
# third-party op restatements
# --------------------------------------------------------------------------------------
def same_pad(in_size: int, k: int, stride: int) -> Tuple[int, int, int]:
    """TF SAME padding: out=ceil(in/stride); extra pad goes AFTER (bottom/right)."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return out, total // 2, total - total // 2


def conv2d(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1) -> Tensor:
    """tf.keras.layers.Conv2D(padding="same"), NHWC x HWIO (+bias)."""
    R, S = w.shape[0], w.shape[1]
    _, pt, pb = same_pad(x.shape[1], R, stride)
    _, pl, pr = same_pad(x.shape[2], S, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn.contiguous(), w.permute(3, 2, 0, 1).contiguous(), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def depthwise_conv2d(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """tf.keras.layers.DepthwiseConv2D((5,5), padding="same"), kernel [5,5,C,1] (decoder.py:130)."""
    C = x.shape[-1]
    k = w.shape[0]
    p = k // 2
    xn = F.pad(x.permute(0, 3, 1, 2), (p, p, p, p))
    y = F.conv2d(xn.contiguous(), w.permute(2, 3, 0, 1).contiguous(), b, groups=C)
    return y.permute(0, 2, 3, 1)


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def elu(x: Tensor) -> Tensor:
    return torch.where(x > 0, x, torch.expm1(torch.clamp(x, max=0.0)))


def upsample_nearest2(x: Tensor, factor: int = 2) -> Tensor:
    """tf.image.resize(method="nearest") by an integer factor == pixel replication (common.py:170-172)."""
    return x.repeat_interleave(factor, dim=1).repeat_interleave(factor, dim=2)


def softclamp5(x: Tensor) -> Tensor:
    """util.py:49-50"""
    return 5.0 * torch.tanh(x / 5.0)


def calculate_log_p(z: Tensor, mu: Tensor, sigma: Tensor) -> Tensor:
    """util.py:39-46"""
    nz = (z - mu) / sigma
    return -0.5 * nz * nz - 0.5 * math.log(2 * math.pi) - torch.log(sigma)


def l2_normalize(x: Tensor) -> Tensor:
    return x * torch.rsqrt(torch.clamp((x * x).sum(), min=1e-12))


# --------------------------------------------------------------------------------------
# stateful layer restatements working on a flat parameter dict
# --------------------------------------------------------------------------------------
class Ctx:
    """Carries parameters, the Keras `training` flag, injected epsilons and a record of tensors."""

    def __init__(self, params: Dict[str, Tensor], training: bool, eps: Optional[List[Tensor]] = None,
                 record: Optional[Dict[str, Tensor]] = None):
        self.p = params
        self.training = training
        self.eps = list(eps) if eps is not None else None
        self.eps_i = 0
        self.record = record
        self.new_stats: Dict[str, Tensor] = {}  # updated BN moving stats / SN u / SN-normalised kernels

    def rec(self, name: str, t: Tensor) -> Tensor:
        if self.record is not None:
            self.record[name] = t
        return t

    def next_eps(self, like: Tensor) -> Tensor:
        if self.eps is None:
            return torch.randn_like(like)
        e = self.eps[self.eps_i]
        self.eps_i += 1
        assert tuple(e.shape) == tuple(like.shape), (e.shape, like.shape)
        return e.to(like.dtype)


def batch_norm(c: Ctx, name: str, x: Tensor) -> Tensor:
    """layers.BatchNormalization(momentum=0.05, epsilon=1e-5), axis -1."""
    g, b = c.p[name + "/gamma"], c.p[name + "/beta"]
    if c.training:
        mean = x.mean(dim=(0, 1, 2))
        var = x.var(dim=(0, 1, 2), unbiased=False)
        n = x.shape[0] * x.shape[1] * x.shape[2]
        with torch.no_grad():
            mm, mv = c.p[name + "/moving_mean"], c.p[name + "/moving_variance"]
            c.new_stats[name + "/moving_mean"] = mm * BN_MOMENTUM + mean * (1 - BN_MOMENTUM)
            c.new_stats[name + "/moving_variance"] = mv * BN_MOMENTUM + var * (n / max(n - 1, 1)) * (1 - BN_MOMENTUM)
    else:
        mean, var = c.p[name + "/moving_mean"], c.p[name + "/moving_variance"]
    return (x - mean) * torch.rsqrt(var + BN_EPS) * g + b


def sn_kernel(c: Ctx, name: str) -> Tensor:
    """tfa.layers.SpectralNormalization(power_iterations=1).normalize_weights:
    one power iteration, kernel overwritten IN PLACE by kernel/sigma when training.  The
    gradient is straight-through to the (already normalised) kernel variable, so the value
    returned here is a fresh leaf that the caller differentiates with respect to."""
    w = c.p[name + "/kernel"]
    if not c.training:
        return w
    key = name + "/kernel"
    if key in c.new_stats:  # already normalised this step (layer reused)
        return c.new_stats[key]
    with torch.no_grad():
        wm = w.reshape(-1, w.shape[-1])
        u = c.p[name + "/u"]
        v = l2_normalize(u @ wm.t())
        u2 = l2_normalize(v @ wm)
        sigma = (v @ wm) @ u2.t()
        wn = (w / sigma).detach()
    wn.requires_grad_(w.requires_grad)
    c.new_stats[key] = wn
    c.new_stats[name + "/u"] = u2
    return wn


def sn_conv(c: Ctx, name: str, x: Tensor, stride: int = 1, use_bias: bool = True) -> Tensor:
    w = sn_kernel(c, name)
    b = c.p[name + "/bias"] if use_bias else None
    return conv2d(x, w, b, stride)


def squeeze_excitation(c: Ctx, name: str, x: Tensor) -> Tensor:
    """common.py:129-142"""
    g = x.mean(dim=(1, 2))
    h = torch.relu(g @ c.p[name + "/dense1/kernel"] + c.p[name + "/dense1/bias"])
    s = torch.sigmoid(h @ c.p[name + "/dense2/kernel"] + c.p[name + "/dense2/bias"])
    return s[:, None, None, :] * x


def rescaler(c: Ctx, name: str, x: Tensor, up: bool, factor: int) -> Tensor:
    """common.py:165-174"""
    x = swish(batch_norm(c, name + "/bn", x))
    if up:
        x = upsample_nearest2(x, factor)
        return sn_conv(c, name + "/conv", x, 1)
    return sn_conv(c, name + "/conv", x, factor)


def encoding_residual_cell(c: Ctx, name: str, x: Tensor) -> Tensor:
    """encoder.py:101-107"""
    t = swish(batch_norm(c, name + "/batch_norm1", x))
    t = sn_conv(c, name + "/conv1", t)
    t = swish(batch_norm(c, name + "/batch_norm2", t))
    t = sn_conv(c, name + "/conv2", t)
    t = squeeze_excitation(c, name + "/se", t)
    return c.rec(name, 0.1 * x + t)


def generative_residual_cell(c: Ctx, name: str, x: Tensor) -> Tensor:
    """decoder.py:138-147"""
    t = batch_norm(c, name + "/batch_norm1", x)
    t = sn_conv(c, name + "/conv1", t)
    t = swish(batch_norm(c, name + "/batch_norm2", t))
    t = depthwise_conv2d(t, c.p[name + "/depth_conv/depthwise_kernel"], c.p[name + "/depth_conv/bias"])
    t = swish(batch_norm(c, name + "/batch_norm3", t))
    t = sn_conv(c, name + "/conv2", t)
    t = batch_norm(c, name + "/batch_norm4", t)
    t = squeeze_excitation(c, name + "/se", t)
    return c.rec(name, 0.1 * x + t)


def encoder_decoder_combiner(c: Ctx, name: str, enc_x: Tensor, dec_x: Tensor) -> Tensor:
    """encoder.py:14-16"""
    return enc_x + sn_conv(c, name + "/decoder_conv", dec_x)


def decoder_sample_combiner(c: Ctx, name: str, x: Tensor, z: Tensor) -> Tensor:
    """decoder.py:114-117"""
    return sn_conv(c, name + "/conv", torch.cat((x, z), dim=3))


@dataclass
class DistributionParams:  # common.py:12-17
    enc_mu: Tensor
    enc_sigma: Tensor
    dec_mu: Tensor
    dec_sigma: Tensor


def sampler_params(c: Ctx, name: str, z_idx: int, prior: Tensor, enc: bool) -> Tuple[Tensor, Tensor]:
    """Sampler.get_params common.py:70-74 (tf.squeeze is a no-op unless a dim is 1)."""
    if enc:
        p = sn_conv(c, f"{name}/enc_sampler/{z_idx}", prior)  # common.py:39-48
    else:
        p = sn_conv(c, f"{name}/dec_sampler/{z_idx}/conv", elu(prior))  # common.py:53-63
    mu, log_sigma = torch.chunk(p, 2, dim=-1)
    return mu, log_sigma


def sampler_call(c: Ctx, name: str, prior: Tensor, z_idx: int, enc_prior: Optional[Tensor] = None):
    """Sampler.call common.py:76-102 with the reparameterisation of :65-68 (epsilon injected)."""
    if enc_prior is None:
        enc_prior = prior
    a, b = sampler_params(c, name, z_idx, enc_prior, enc=True)
    if z_idx == 0:
        enc_mu = softclamp5(a)
        enc_sigma = torch.exp(softclamp5(b)) + 1e-2
        z = enc_mu + c.next_eps(enc_mu) * enc_sigma
        return z, DistributionParams(enc_mu, enc_sigma, torch.zeros_like(enc_mu), torch.ones_like(enc_sigma))
    cm, cs = sampler_params(c, name, z_idx, prior, enc=False)
    dec_mu = softclamp5(cm)
    dec_sigma = torch.exp(softclamp5(cs)) + 1e-2
    enc_mu = softclamp5(a + cm)
    enc_sigma = torch.exp(softclamp5(cs + b)) + 1e-2
    z = enc_mu + c.next_eps(enc_mu) * enc_sigma
    return z, DistributionParams(enc_mu, enc_sigma, dec_mu, dec_sigma)


# --------------------------------------------------------------------------------------
# model structure (shared by parameter creation and the forward walk)
# --------------------------------------------------------------------------------------
@dataclass
class Structure:
    """Static description of the layer graph, derived exactly as the reference constructors do."""
    cfg: NVAEConfig
    pre: List[dict] = field(default_factory=list)
    enc: List[dict] = field(default_factory=list)
    dec: List[dict] = field(default_factory=list)
    post: List[dict] = field(default_factory=list)
    enc_final_channels: int = 0
    z0_hw: int = 0
    samplers: List[dict] = field(default_factory=list)  # per z_idx: enc in-ch, dec in-ch


def build_structure(cfg: NVAEConfig) -> Structure:
    s = Structure(cfg)
    sf = cfg.scale_factor
    ce = cfg.n_encoder_channels
    # Preprocess (preprocess.py:19-36)
    mult = 1
    cin = ce
    hw = cfg.image_size
    for _ in range(cfg.n_preprocess_blocks):
        for _ in range(cfg.n_preprocess_cells - 1):
            s.pre.append(dict(kind="bnswishconv", cin=cin, cout=mult * ce, stride=1))
            cin = mult * ce
        s.pre.append(dict(kind="bnswishconv", cin=cin, cout=mult * ce * sf, stride=2))
        cin = mult * ce * sf
        mult *= sf
        hw //= sf
    # Encoder (encoder.py:34-68)
    n_scales = cfg.n_latent_scales
    for scale in range(n_scales):
        n_groups = cfg.n_groups_per_scale[scale]
        for g in range(n_groups):
            ch = ce * mult
            s.enc.append(dict(kind="cells", ch=ch, n=cfg.res_cells_per_group, hw=hw))
            if not (scale == n_scales - 1 and g == n_groups - 1):
                s.enc.append(dict(kind="combiner", ch=ch, hw=hw))
        if scale < n_scales - 1:
            s.enc.append(dict(kind="rescaler", cin=ce * mult, cout=ce * mult * sf, up=False))
            mult *= sf
            hw //= sf
    s.enc_final_channels = ce * mult
    s.z0_hw = hw
    # Decoder (decoder.py:24-62); groups per scale reversed by models.py:70
    cd = cfg.n_decoder_channels
    rev = list(reversed(cfg.n_groups_per_scale))
    dmult = float(mult)
    cin = cd  # channels of h (decoder.py:57-62)
    for scale in range(n_scales):
        for g in range(rev[scale]):
            ch = int(cd * dmult)
            if not (scale == 0 and g == 0):
                s.dec.append(dict(kind="cells", ch=ch, n=cfg.res_cells_per_group, hw=hw))
                s.samplers.append(dict(enc_cin=ch, dec_cin=ch, hw=hw))
            else:
                s.samplers.append(dict(enc_cin=s.enc_final_channels, dec_cin=None, hw=hw))
            s.dec.append(dict(kind="dsc", cin=cin + cfg.n_latent_per_group, cout=ch, hw=hw))
            cin = ch
        if scale < n_scales - 1:
            s.dec.append(dict(kind="rescaler", cin=cin, cout=int(cd * dmult / sf), up=True))
            cin = int(cd * dmult / sf)
            dmult /= sf
            hw *= sf
    # Postprocess (postprocess.py:13-31)
    pmult = dmult
    for _ in range(cfg.n_postprocess_blocks):
        pmult /= sf
        ch = int(cd * pmult)
        for cell_idx in range(cfg.n_post_process_cells):
            s.post.append(dict(kind="postcell", cin=cin, ch=ch, up=(cell_idx == 0)))
            cin = ch
    s.post.append(dict(kind="final", cin=cin))
    return s


def _glorot(rng: np.random.Generator, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape)


class ParamBuilder:
    """Creates Keras-default initialised parameters keyed by attribute path."""

    def __init__(self, seed: int, jitter: float = 0.0):
        self.rng = np.random.default_rng(seed)
        self.jitter = jitter
        self.params: Dict[str, np.ndarray] = {}
        self.trainable: List[str] = []
        self.bn_in_loss: List[str] = []  # BN layers reached by calculate_bn_loss (models.py:252-267)

    def _add(self, name, arr, trainable=True):
        self.params[name] = np.asarray(arr, dtype=np.float64)
        if trainable:
            self.trainable.append(name)

    def conv(self, name, k, cin, cout, bias=True, sn=True):
        self._add(name + "/kernel", _glorot(self.rng, (k, k, cin, cout), k * k * cin, k * k * cout))
        if bias:
            self._add(name + "/bias", self.rng.normal(0, self.jitter, size=(cout,)) if self.jitter else np.zeros(cout))
        if sn:
            u = np.clip(self.rng.normal(0, 0.02, size=(1, cout)), -0.04, 0.04)  # TruncatedNormal(stddev=0.02)
            self._add(name + "/u", u, trainable=False)

    def depthwise(self, name, c):
        # Keras DepthwiseConv2D glorot_uniform on shape [5,5,C,1]: fan_in=25*C, fan_out=25*1
        self._add(name + "/depthwise_kernel", _glorot(self.rng, (5, 5, c, 1), 25 * c, 25))
        self._add(name + "/bias", self.rng.normal(0, self.jitter, size=(c,)) if self.jitter else np.zeros(c))

    def bn(self, name, c, in_loss=False):
        j = self.jitter
        self._add(name + "/gamma", 1.0 + (self.rng.normal(0, j, size=(c,)) if j else np.zeros(c)))
        self._add(name + "/beta", self.rng.normal(0, j, size=(c,)) if j else np.zeros(c))
        self._add(name + "/moving_mean", self.rng.normal(0, j, size=(c,)) if j else np.zeros(c), trainable=False)
        self._add(name + "/moving_variance", 1.0 + (np.abs(self.rng.normal(0, j, size=(c,))) if j else np.zeros(c)),
                  trainable=False)
        if in_loss:
            self.bn_in_loss.append(name)

    def dense(self, name, cin, cout):
        self._add(name + "/kernel", _glorot(self.rng, (cin, cout), cin, cout))
        self._add(name + "/bias", self.rng.normal(0, self.jitter, size=(cout,)) if self.jitter else np.zeros(cout))

    def se(self, name, c):
        hid = int(max(c / 16, 4))  # common.py:125
        self.dense(name + "/dense1", c, hid)
        self.dense(name + "/dense2", hid, c)


def build_params(cfg: NVAEConfig, seed: int = 1, jitter: float = 0.0):
    """Returns (params: name->float64 ndarray, trainable names, bn-in-loss layer names, Structure)."""
    s = build_structure(cfg)
    pb = ParamBuilder(seed, jitter)
    lat2 = 2 * cfg.n_latent_per_group
    # --- preprocess
    pb.conv("preprocess/stem", 3, cfg.image_channels, cfg.n_encoder_channels)
    for i, d in enumerate(s.pre):
        n = f"preprocess/cells/{i}"
        cin = d["cin"]
        for j in range(2):
            pb.bn(f"{n}/nodes/{j}/bn", cin)
            pb.conv(f"{n}/nodes/{j}/conv", 3, cin, d["cout"])
            cin = d["cout"]
        pb.se(f"{n}/se", d["cout"])
        if d["stride"] == 2:
            q = d["cout"] // 4
            for k in range(3):
                pb.conv(f"{n}/skip/conv{k + 1}", 1, d["cin"], q)
            pb.conv(f"{n}/skip/conv4", 1, d["cin"], d["cout"] - 3 * q)
    # --- encoder
    for i, d in enumerate(s.enc):
        n = f"encoder/groups/{i}"
        if d["kind"] == "cells":
            for k in range(d["n"]):
                cn = f"{n}/cells/{k}"
                pb.bn(cn + "/batch_norm1", d["ch"], in_loss=True)
                pb.conv(cn + "/conv1", 3, d["ch"], d["ch"])
                pb.bn(cn + "/batch_norm2", d["ch"], in_loss=True)
                pb.conv(cn + "/conv2", 3, d["ch"], d["ch"])
                pb.se(cn + "/se", d["ch"])
        elif d["kind"] == "combiner":
            pb.conv(n + "/decoder_conv", 1, d["ch"], d["ch"])
        else:
            pb.bn(n + "/bn", d["cin"], in_loss=True)
            pb.conv(n + "/conv", 3, d["cin"], d["cout"])
    pb.conv("encoder/final_enc/conv", 1, s.enc_final_channels, s.enc_final_channels)
    # --- decoder
    for zi, sd in enumerate(s.samplers):
        pb.conv(f"decoder/sampler/enc_sampler/{zi}", 3, sd["enc_cin"], lat2)
        if sd["dec_cin"] is not None:
            pb.conv(f"decoder/sampler/dec_sampler/{zi}/conv", 1, sd["dec_cin"], lat2)
    for i, d in enumerate(s.dec):
        n = f"decoder/groups/{i}"
        if d["kind"] == "cells":
            for k in range(d["n"]):
                cn = f"{n}/cells/{k}"
                ch = d["ch"]
                pb.bn(cn + "/batch_norm1", ch, in_loss=True)
                pb.conv(cn + "/conv1", 1, ch, 6 * ch)
                pb.bn(cn + "/batch_norm2", 6 * ch, in_loss=True)
                pb.depthwise(cn + "/depth_conv", 6 * ch)
                pb.bn(cn + "/batch_norm3", 6 * ch, in_loss=True)
                pb.conv(cn + "/conv2", 1, 6 * ch, ch)
                pb.bn(cn + "/batch_norm4", ch, in_loss=True)
                pb.se(cn + "/se", ch)
        elif d["kind"] == "dsc":
            pb.conv(n + "/conv", 1, d["cin"], d["cout"])
        else:
            pb.bn(n + "/bn", d["cin"], in_loss=True)
            pb.conv(n + "/conv", 3, d["cin"], d["cout"])
    pb._add("decoder/h", pb.rng.uniform(0, 1, size=(s.z0_hw, s.z0_hw, cfg.n_decoder_channels)))  # decoder.py:60-62
    # --- postprocess
    for i, d in enumerate(s.post):
        n = f"postprocess/cells/{i}"
        if d["kind"] == "final":
            pb.conv("postprocess/final", 3, d["cin"], 1)
            continue
        ch, cin = d["ch"], d["cin"]
        if d["up"]:
            pb.bn(n + "/skip/bn", cin)
            pb.conv(n + "/skip/conv", 3, cin, ch)
            pb.bn(n + "/node/rescaler/bn", cin)
            pb.conv(n + "/node/rescaler/conv", 3, cin, ch)
        pb.bn(n + "/node/bn0", ch)
        pb.conv(n + "/node/cbs1/conv", 1, ch, 6 * ch, bias=False)
        pb.bn(n + "/node/cbs1/bn", 6 * ch)
        pb.conv(n + "/node/cbs2/conv", 5, 6 * ch, 6 * ch, bias=False)
        pb.bn(n + "/node/cbs2/bn", 6 * ch)
        pb.conv(n + "/node/conv3", 1, 6 * ch, ch, bias=False)
        pb.bn(n + "/node/bn1", ch)
        pb.se(n + "/node/se", ch)
    return pb.params, pb.trainable, pb.bn_in_loss, s


# --------------------------------------------------------------------------------------
# forward walk (models.py:89-98)
# --------------------------------------------------------------------------------------
def bn_swish_conv(c: Ctx, name: str, d: dict, x: Tensor) -> Tensor:
    """preprocess.py:77-107 (BNSwishConv) with SkipScaler preprocess.py:42-74"""
    t = x
    for j in range(2):
        t = swish(batch_norm(c, f"{name}/nodes/{j}/bn", t))
        t = sn_conv(c, f"{name}/nodes/{j}/conv", t, d["stride"] if j == 0 else 1)
    t = squeeze_excitation(c, name + "/se", t)
    if d["stride"] == 1:
        skipped = x
    else:
        o = swish(x)
        skipped = torch.cat((
            sn_conv(c, name + "/skip/conv1", o, 2),
            sn_conv(c, name + "/skip/conv2", o[:, 1:, 1:, :], 2),
            sn_conv(c, name + "/skip/conv3", o[:, :, 1:, :], 2),
            sn_conv(c, name + "/skip/conv4", o[:, 1:, :, :], 2)), dim=3)
    return c.rec(name, skipped + 0.1 * t)


def preprocess(c: Ctx, s: Structure, x: Tensor) -> Tensor:
    """preprocess.py:37-39"""
    t = sn_conv(c, "preprocess/stem", 2 * x - 1)
    for i, d in enumerate(s.pre):
        t = bn_swish_conv(c, f"preprocess/cells/{i}", d, t)
    return t


def encoder(c: Ctx, s: Structure, x: Tensor):
    """encoder.py:70-83: returns the deferred combiners [(name, enc_x)] and the final encoding."""
    combiners = []
    for i, d in enumerate(s.enc):
        n = f"encoder/groups/{i}"
        if d["kind"] == "combiner":
            combiners.append((n, x))
        elif d["kind"] == "cells":
            for k in range(d["n"]):
                x = encoding_residual_cell(c, f"{n}/cells/{k}", x)
        else:
            x = c.rec(n, rescaler(c, n, x, up=False, factor=s.cfg.scale_factor))
    final = elu(sn_conv(c, "encoder/final_enc/conv", elu(x)))
    return combiners, c.rec("encoder/final", final)


def decoder(c: Ctx, s: Structure, prior: Tensor, combiners, nll: bool = False):
    """decoder.py:64-104"""
    z_params: List[DistributionParams] = []
    log_p = prior.new_zeros(prior.shape[0])
    log_q = prior.new_zeros(prior.shape[0])
    z0, params = sampler_call(c, "decoder/sampler", prior, 0)
    c.rec("z/0", z0)
    z_params.append(params)
    if nll:
        log_q = log_q + calculate_log_p(z0, params.enc_mu, params.enc_sigma).sum(dim=(1, 2, 3))
        log_p = log_p + calculate_log_p(z0, params.dec_mu, params.dec_sigma).sum(dim=(1, 2, 3))
    h = c.p["decoder/h"].unsqueeze(0).expand(z0.shape[0], -1, -1, -1)
    x = decoder_sample_combiner(c, "decoder/groups/0", h, z0)
    ci = 0
    for i, d in enumerate(s.dec[1:], start=1):
        n = f"decoder/groups/{i}"
        if d["kind"] == "dsc":
            cn, enc_x = combiners[ci]
            enc_prior = encoder_decoder_combiner(c, cn, enc_x, x)
            z, params = sampler_call(c, "decoder/sampler", x, ci + 1, enc_prior)
            c.rec(f"z/{ci + 1}", z)
            if nll:
                log_q = log_q + calculate_log_p(z, params.enc_mu, params.enc_sigma).sum(dim=(1, 2, 3))
                log_p = log_p + calculate_log_p(z, params.dec_mu, params.dec_sigma).sum(dim=(1, 2, 3))
            z_params.append(params)
            x = c.rec(n, decoder_sample_combiner(c, n, x, z))
            ci += 1
        elif d["kind"] == "cells":
            for k in range(d["n"]):
                x = generative_residual_cell(c, f"{n}/cells/{k}", x)
        else:
            x = c.rec(n, rescaler(c, n, x, up=True, factor=s.cfg.scale_factor))
    return x, z_params, log_p, log_q


def postprocess(c: Ctx, s: Structure, x: Tensor) -> Tensor:
    """postprocess.py:33-34, :57-58, :66-88, :94-108"""
    sf = s.cfg.scale_factor
    for i, d in enumerate(s.post):
        n = f"postprocess/cells/{i}"
        if d["kind"] == "final":
            return sn_conv(c, "postprocess/final", elu(x))
        t = x
        if d["up"]:
            skip = rescaler(c, n + "/skip", x, up=True, factor=sf)
            t = rescaler(c, n + "/node/rescaler", t, up=True, factor=sf)
        else:
            skip = x
        t = batch_norm(c, n + "/node/bn0", t)
        t = swish(batch_norm(c, n + "/node/cbs1/bn", sn_conv(c, n + "/node/cbs1/conv", t, use_bias=False)))
        t = swish(batch_norm(c, n + "/node/cbs2/bn", sn_conv(c, n + "/node/cbs2/conv", t, use_bias=False)))
        t = sn_conv(c, n + "/node/conv3", t, use_bias=False)
        t = batch_norm(c, n + "/node/bn1", t)
        t = squeeze_excitation(c, n + "/node/se", t)
        x = c.rec(n, skip + 0.1 * t)
    raise AssertionError("postprocess must end with the final conv")


def nvae_call(c: Ctx, s: Structure, inputs: Tensor, nll: bool = False):
    """NVAE.call models.py:89-98"""
    x = c.rec("preprocess", preprocess(c, s, inputs))
    combiners, final_x = encoder(c, s, x)
    combiners = list(reversed(combiners))  # models.py:93
    dec, z_params, log_p, log_q = decoder(c, s, final_x, combiners, nll=nll)
    c.rec("decoder", dec)
    logits = c.rec("logits", postprocess(c, s, dec))
    return logits, z_params, log_p, log_q


# --------------------------------------------------------------------------------------
# losses (models.py:191-267)
# --------------------------------------------------------------------------------------
def kl_per_group(z_params: List[DistributionParams]) -> Tensor:
    """models.py:197-201 -> [G, B]"""
    out = []
    for g in z_params:
        t1 = (g.enc_mu - g.dec_mu) / g.dec_sigma
        t2 = g.enc_sigma / g.dec_sigma
        kl = 0.5 * (t1 * t1 + t2 * t2) - 0.5 - torch.log(t2)
        out.append(kl.sum(dim=(1, 2, 3)))
    return torch.stack(out, 0)


def kl_alphas(cfg: NVAEConfig) -> np.ndarray:
    """models.py:227-237"""
    gps = cfg.n_groups_per_scale
    ns = len(gps)
    coeffs = []
    for i in range(ns):
        n = gps[ns - i - 1]
        coeffs.append(np.square(2 ** i) / n * np.ones(n))
    coeffs = np.concatenate(coeffs)
    return coeffs / coeffs.min()


def calculate_kl_loss(cfg: NVAEConfig, z_params, balancing: bool) -> Tuple[Tensor, Tensor]:
    """models.py:191-223 -> (loss [B], kl_all [G,B])"""
    kl_all = kl_per_group(z_params)
    if balancing:
        alphas = torch.as_tensor(kl_alphas(cfg), dtype=kl_all.dtype)
        coeff = kl_all.abs().mean(dim=1) + 0.01
        total = coeff.sum()
        coeff = coeff / alphas * total
        coeff = coeff / coeff.mean()
        loss = (kl_all * coeff.detach()[:, None]).sum(dim=0)
    else:
        loss = kl_all.sum(dim=0)
    return loss, kl_all


def calculate_recon_loss(inputs: Tensor, logits: Tensor, crop_output: bool = False) -> Tensor:
    """models.py:242-250; TFP Bernoulli(logits).log_prob(x) = x*l - softplus(l)."""
    if crop_output:
        inputs = inputs[:, 2:30, 2:30, :]
        logits = logits[:, 2:30, 2:30, :]
    log_probs = inputs * logits - F.softplus(logits)
    return -log_probs.sum(dim=(1, 2, 3))


def calculate_bn_loss(c: Ctx, bn_in_loss: List[str], sr_lambda: float) -> Tensor:
    """models.py:252-267: sr_lambda * sum over the 88 encoder/decoder-group BN layers of max|gamma|."""
    tot = 0.0
    for n in bn_in_loss:
        tot = tot + c.p[n + "/gamma"].abs().max()
    return sr_lambda * tot


def beta_schedule(cfg: NVAEConfig, steps: int, epoch: int = 0) -> float:
    """models.py:121-122"""
    m = steps if cfg.step_based_warmup else epoch
    return min(m / (0.3 * cfg.n_total_iterations), 1)


def train_step_loss(cfg: NVAEConfig, s: Structure, params: Dict[str, Tensor], bn_in_loss: List[str],
                    data: Tensor, eps: List[Tensor], steps: int, training: bool = True,
                    record: Optional[Dict[str, Tensor]] = None):
    """The differentiable part of NVAE.train_step (models.py:116-126).
    Returns (dict of losses as in models.py:130-135, Ctx)."""
    c = Ctx(params, training, eps, record)
    logits, z_params, _, _ = nvae_call(c, s, data)
    recon = calculate_recon_loss(data, logits)
    bn_loss = calculate_bn_loss(c, bn_in_loss, cfg.sr_lambda)
    beta = beta_schedule(cfg, steps)
    kl, kl_all = calculate_kl_loss(cfg, z_params, beta < 1)
    kl_loss = beta * kl
    loss = (recon + kl_loss).mean()
    total = loss + bn_loss
    return dict(loss=total, reconstruction_loss=recon, kl_loss=kl_loss, bn_loss=bn_loss,
                kl_all=kl_all, logits=logits, z_params=z_params), c


def grads_wrt_trainables(total: Tensor, c: Ctx, trainable: List[str]) -> Dict[str, Tensor]:
    """tape.gradient(total_loss, trainable_weights) (models.py:127).  SN-wrapped kernels are
    differentiated at their normalised value (straight-through)."""
    leaves = [c.new_stats.get(n, c.p[n]) if n.endswith("/kernel") else c.p[n] for n in trainable]
    gs = torch.autograd.grad(total, leaves, allow_unused=True)
    return {n: (g if g is not None else torch.zeros_like(l)) for n, g, l in zip(trainable, gs, leaves)}


# --------------------------------------------------------------------------------------
# optimizer (train.py:128-131)
# --------------------------------------------------------------------------------------
def cosine_decay_lr(step: int, decay_steps: int, lr0: float = 1e-3) -> float:
    t = min(step, decay_steps)
    return lr0 * 0.5 * (1.0 + math.cos(math.pi * t / decay_steps))


def adamax_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, t: int, lr: float,
                  b1: float = 0.9, b2: float = 0.999, eps: float = 1e-7):
    """tf.keras.optimizers.Adamax dense update; t is the 1-based iteration."""
    m = b1 * m + (1 - b1) * g
    v = torch.maximum(b2 * v, g.abs())
    p = p - (lr / (1 - b1 ** t)) * m / (v + eps)
    return p, m, v


# --------------------------------------------------------------------------------------
# ancestral sampling (models.py:137-178), inference-mode BN, no SN
# --------------------------------------------------------------------------------------
def sample(cfg: NVAEConfig, s: Structure, params: Dict[str, Tensor], n_samples: int, temperature: float,
           eps: List[Tensor]):
    c = Ctx(params, False, eps)
    h = c.p["decoder/h"]
    st = h.unsqueeze(0).expand(n_samples, -1, -1, -1)
    z0_shape = (n_samples, s.z0_hw, s.z0_hw, cfg.n_latent_per_group)
    mu = softclamp5(h.new_zeros(z0_shape))
    sigma = torch.exp(softclamp5(h.new_zeros(z0_shape))) + 1e-2
    if temperature != 1.0:
        sigma = sigma * temperature  # only z0 is tempered (models.py:143-144)
    z = mu + c.next_eps(mu) * sigma
    di = 0
    for i, d in enumerate(s.dec):
        n = f"decoder/groups/{i}"
        if d["kind"] == "dsc":
            if di > 0:
                m, ls = sampler_params(c, "decoder/sampler", di, st, enc=False)
                mu = softclamp5(m)
                sigma = torch.exp(softclamp5(ls)) + 1e-2
                z = mu + c.next_eps(mu) * sigma
            st = decoder_sample_combiner(c, n, st, z)
            di += 1
        elif d["kind"] == "cells":
            for k in range(d["n"]):
                st = generative_residual_cell(c, f"{n}/cells/{k}", st)
        else:
            st = rescaler(c, n, st, up=True, factor=cfg.scale_factor)
    logits = postprocess(c, s, st)
    return torch.sigmoid(logits), logits


# --------------------------------------------------------------------------------------
# helpers for tests / bench
# --------------------------------------------------------------------------------------
def to_torch(params: Dict[str, np.ndarray], trainable: List[str], dtype=torch.float64) -> Dict[str, Tensor]:
    tset = set(trainable)
    out = {}
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v), dtype=dtype).clone()
        if k in tset:
            t.requires_grad_(True)
        out[k] = t
    return out


def make_eps(s: Structure, batch: int, seed: int = 1, dtype=torch.float64) -> List[Tensor]:
    """One epsilon tensor per latent group in top-down order."""
    rng = np.random.default_rng(seed)
    L = s.cfg.n_latent_per_group
    return [torch.as_tensor(rng.standard_normal((batch, sd["hw"], sd["hw"], L)), dtype=dtype) for sd in s.samplers]


def make_images(cfg: NVAEConfig, batch: int, seed: int = 1, dtype=torch.float64) -> Tensor:
    """28x28 Bernoulli(p=0.13) {0,1} images zero-padded to 32x32 (datasets.py:12)."""
    rng = np.random.default_rng(seed)
    pad = 2 if cfg.image_size == 32 else 0
    inner = cfg.image_size - 2 * pad
    x = (rng.random((batch, inner, inner, cfg.image_channels)) < 0.13).astype(np.float64)
    x = np.pad(x, ((0, 0), (pad, pad), (pad, pad), (0, 0)))
    return torch.as_tensor(x, dtype=dtype)
