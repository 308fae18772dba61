"""Drop-in for the reference's common.py: DistributionParams, Sampler, RescaleType,
SqueezeExcitation, Rescaler -- same class names, constructor arguments and call signatures
(common.py:12-174), driving libnvae_b200 kernels instead of TensorFlow ops.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum, auto
from typing import List, Optional, Sequence, Tuple

import torch

from . import runtime as R
from ._lib import NVAE_ACT_ELU, NVAE_ACT_NONE, NVAE_ACT_SWISH
from .layers import BatchNormalization, Conv2D, Dense, Layer, SpectralNormalization
from .runtime import DeviceTensor


@dataclass
class DistributionParams:
    """common.py:12-17.  The four tensors are views of one [4,B,h,w,L] buffer the latent kernel fills;
    `kl` is the per-sample KL row of this group ([B], models.py:197-201 fused into the same kernel)."""
    enc_mu: torch.Tensor
    enc_sigma: torch.Tensor
    dec_mu: torch.Tensor
    dec_sigma: torch.Tensor
    kl: Optional[torch.Tensor] = None
    group: int = 0


class _DecSampler(Layer):
    """Sequential([ELU(), SpectralNormalization(Conv2D(2L, (1,1)))]) of common.py:53-63."""

    def __init__(self, latents2: int, in_channels: int, name: str):
        super().__init__(name)
        with self.rt.scope(name):
            self.conv = SpectralNormalization(Conv2D(latents2, kernel_size=(1, 1), in_channels=in_channels))

    def __call__(self, x: DeviceTensor, training: bool = False) -> DeviceTensor:
        return self.conv(R.bn_act(self.rt, x, None, NVAE_ACT_ELU, False), training)


class Sampler(Layer):
    def __init__(self, n_latent_scales, n_groups_per_scale, n_latent_per_group, scale_factor, *,
                 group_channels: Sequence[int], name: str = "sampler", **kwargs) -> None:
        """`group_channels[z_idx]` = channels of the feature map group z_idx samples from (Keras infers
        them at build time; the arena needs them up front)."""
        super().__init__(name)
        self.enc_sampler: List[SpectralNormalization] = []
        self.dec_sampler: List[Optional[_DecSampler]] = []
        self.n_latent_scales = n_latent_scales
        self.n_groups_per_scale = n_groups_per_scale
        self.n_latent_per_group = n_latent_per_group
        idx = 0
        with self.rt.scope(name):
            for scale in range(self.n_latent_scales):
                n_groups = self.n_groups_per_scale[scale]
                for group in range(n_groups):
                    with self.rt.scope("enc_sampler"):
                        self.enc_sampler.append(SpectralNormalization(
                            Conv2D(2 * self.n_latent_per_group, kernel_size=(3, 3), padding="same",
                                   in_channels=group_channels[idx], name=str(idx))))
                    if scale == 0 and group == 0:
                        self.dec_sampler.append(None)  # dummy to maintain indexing (common.py:50-52)
                    else:
                        with self.rt.scope("dec_sampler"):
                            self.dec_sampler.append(
                                _DecSampler(2 * self.n_latent_per_group, group_channels[idx], name=str(idx)))
                    idx += 1
        # per-forward buffers set by Decoder.call / NVAE.sample
        self.kl_all: Optional[torch.Tensor] = None     # [G,B]
        self.kl_weight: Optional[torch.Tensor] = None  # [G] d(total)/d(kl[g,b]), written by the loss kernel

    @property
    def n_groups(self) -> int:
        return len(self.enc_sampler)

    def begin(self, batch: int) -> None:
        G = self.n_groups
        self.kl_all = self.rt.zeros(G, batch)
        self.kl_weight = self.rt.zeros(G)

    def sample(self, mu: torch.Tensor, sigma: torch.Tensor, sigma_scale: float = 1.0) -> torch.Tensor:
        """Reparametrisation trick (common.py:65-68) on materialised mu/sigma (NVAE.sample's draws)."""
        rt = self.rt
        eps = rt.next_eps(tuple(mu.shape))
        z = rt.empty(*mu.shape)
        rt.lib.reparam(mu.data_ptr(), sigma.data_ptr(), eps.data_ptr(), sigma_scale, z.data_ptr(), z.numel(), rt.stream)
        return z

    def get_params(self, sampler, z_idx, prior: DeviceTensor, training: bool = False) -> DeviceTensor:
        """common.py:70-74.  Returns the raw [B,h,w,2L] conv output; the (mu | log_sigma) split is
        an addressing convention of the latent kernels (tf.split/tf.squeeze launch nothing here)."""
        return sampler[z_idx](prior, training)

    def __call__(self, prior: DeviceTensor, z_idx: int, enc_prior: Optional[DeviceTensor] = None,
                 training: bool = False, log_q: Optional[torch.Tensor] = None,
                 log_p: Optional[torch.Tensor] = None, want_params: bool = True
                 ) -> Tuple[DeviceTensor, DistributionParams]:
        """common.py:76-102 + the per-group KL row sum of models.py:197-201, one fused launch."""
        rt = self.rt
        if enc_prior is None:
            enc_prior = prior
        enc_p = self.get_params(self.enc_sampler, z_idx, enc_prior, training)
        dec_p = None if z_idx == 0 else self.get_params(self.dec_sampler, z_idx, prior, training)
        B, h, w, L2 = enc_p.shape
        L = L2 // 2
        if self.kl_all is None or self.kl_all.shape[1] != B:
            self.begin(B)
        eps = rt.next_eps((B, h, w, L))
        z = DeviceTensor(rt.empty(B, h, w, L))
        dist = rt.empty(4, B, h, w, L) if want_params else None
        kl = self.kl_all[z_idx]
        rt.lib.latent_fwd(enc_p.ptr(), dec_p.ptr() if dec_p is not None else None, eps.data_ptr(), B, h * w, L, z.ptr(),
                          kl.data_ptr(), log_q.data_ptr() if log_q is not None else None,
                          log_p.data_ptr() if log_p is not None else None,
                          dist.data_ptr() if dist is not None else None, rt.stream)
        if rt.tape is not None:
            klw = self.kl_weight[z_idx:z_idx + 1]

            def bwd():
                d_enc = rt.empty(B, h, w, L2)
                d_dec = rt.empty(B, h, w, L2) if dec_p is not None else None
                rt.lib.latent_bwd(enc_p.ptr(), dec_p.ptr() if dec_p is not None else None, eps.data_ptr(),
                                  z.grad.data_ptr() if z.grad is not None else None, klw.data_ptr(), B, h * w, L,
                                  d_enc.data_ptr(), d_dec.data_ptr() if d_dec is not None else None, rt.stream)
                rt.add_grad(enc_p, d_enc)
                if dec_p is not None:
                    rt.add_grad(dec_p, d_dec)
                z.grad = None
            rt.record(bwd)
        if dist is not None:
            params = DistributionParams(dist[0], dist[1], dist[2], dist[3], kl=kl, group=z_idx)
        else:
            params = DistributionParams(None, None, None, None, kl=kl, group=z_idx)
        return z, params

    call = __call__


class RescaleType(Enum):
    UP = auto()
    DOWN = auto()


class SqueezeExcitation(Layer):
    """Squeeze and Excitation block (Hu et al. 2019), common.py:110-142."""

    def __init__(self, ratio=16, *, channels: int, name: str = "se", **kwargs) -> None:
        super().__init__(name)
        self.ratio = ratio
        c = channels
        num_hidden = max(c / self.ratio, 4)  # common.py:125 (Keras casts the float to int)
        with self.rt.scope(name):
            self.dense1 = Dense(units=num_hidden, in_features=c, name="dense1")
            self.dense2 = Dense(units=c, in_features=int(num_hidden), name="dense2")

    def __call__(self, inputs: DeviceTensor, training: bool = False) -> DeviceTensor:
        """x * sigmoid(dense2(relu(dense1(gap(x)))))  == the fused kernel with alpha=0, beta=1."""
        return R.se_residual(self.rt, inputs, None, inputs, self, 0.0, 1.0, training)

    def fused(self, t: DeviceTensor, xres: DeviceTensor, alpha: float, beta: float, bn=None,
              training: bool = False) -> DeviceTensor:
        """alpha*xres + beta*SE(BN(t)) in two launches (the residual tails encoder.py:107, decoder.py:147,
        preprocess.py:107, postprocess.py:58)."""
        return R.se_residual(self.rt, t, bn, xres, self, alpha, beta, training)

    call = __call__


class Rescaler(Layer):
    def __init__(self, n_channels, scale_factor, rescale_type, *, in_channels: int, name: str = "rescaler",
                 in_bn_loss: bool = False, **kwargs) -> None:
        super().__init__(name)
        self.mode = rescale_type
        self.factor = scale_factor
        if scale_factor != 2:
            raise ValueError("this path implements the reference's scale_factor=2 (train.py:208-214)")
        with self.rt.scope(name):
            self.bn = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=in_channels, name="bn",
                                         in_bn_loss=in_bn_loss)
            if rescale_type == RescaleType.UP:
                self.conv = SpectralNormalization(
                    Conv2D(n_channels, (3, 3), strides=(1, 1), padding="same", in_channels=in_channels))
            elif rescale_type == RescaleType.DOWN:
                self.conv = SpectralNormalization(
                    Conv2D(n_channels, (3, 3), strides=(self.factor, self.factor), padding="same",
                           in_channels=in_channels))

    def __call__(self, input: DeviceTensor, training: bool = False) -> DeviceTensor:
        # BN -> swish -> [nearest x2] in one apply launch (common.py:166-172), then the SN conv
        x = R.bn_act(self.rt, input, self.bn, NVAE_ACT_SWISH, training, upsample=self.mode == RescaleType.UP)
        return self.conv(x, training)

    call = __call__
