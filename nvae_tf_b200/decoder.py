"""Drop-in for the reference's decoder.py: Decoder, DecoderSampleCombiner, GenerativeResidualCell
(decoder.py:9-147) on libnvae_b200 kernels."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from . import runtime as R
from ._lib import NVAE_ACT_NONE, NVAE_ACT_SWISH
from .common import RescaleType, Rescaler, Sampler, SqueezeExcitation
from .encoder import _Sequential
from .layers import BatchNormalization, Conv2D, DepthwiseConv2D, Layer, SpectralNormalization
from .runtime import DeviceTensor


class Decoder(Layer):
    def __init__(self, n_decoder_channels, n_latent_per_group: int, res_cells_per_group, n_latent_scales: int,
                 n_groups_per_scale: List[int], mult: int, scale_factor: int, input_shape, *, name: str = "decoder",
                 **kwargs):
        super().__init__(name)
        rt = self.rt
        self.groups = []
        self.n_decoder_channels = n_decoder_channels
        group_channels = []
        in_channels = n_decoder_channels  # channels of h (decoder.py:57-62)
        with rt.scope(name):
            with rt.scope("groups"):
                for scale in range(n_latent_scales):
                    n_groups = n_groups_per_scale[scale]
                    for group in range(n_groups):
                        output_channels = int(n_decoder_channels * mult)  # Keras casts (decoder.py:35)
                        if not (scale == 0 and group == 0):
                            g = _Sequential(str(len(self.groups)))
                            with rt.scope(g.name), rt.scope("cells"):
                                for k in range(res_cells_per_group):
                                    g.add(GenerativeResidualCell(output_channels, name=str(k)))
                            self.groups.append(g)
                            group_channels.append(output_channels)
                        else:
                            group_channels.append(int(input_shape[3]))  # z0 is sampled from the encoder output
                        self.groups.append(DecoderSampleCombiner(
                            output_channels, in_channels=in_channels + n_latent_per_group,
                            name=str(len(self.groups))))
                        in_channels = output_channels
                    if scale < n_latent_scales - 1:
                        output_channels = int(n_decoder_channels * mult / scale_factor)
                        self.groups.append(Rescaler(output_channels, scale_factor=scale_factor,
                                                    rescale_type=RescaleType.UP, in_channels=in_channels,
                                                    name=str(len(self.groups)), in_bn_loss=True))
                        in_channels = output_channels
                        mult /= scale_factor
            self.sampler = Sampler(n_latent_scales=n_latent_scales, n_groups_per_scale=n_groups_per_scale,
                                   n_latent_per_group=n_latent_per_group, scale_factor=scale_factor,
                                   group_channels=group_channels)
            self.mult = mult
            self.out_channels = in_channels
            self.z0_shape = [int(input_shape[1]), int(input_shape[2]), n_latent_per_group]
            h_shape = [int(input_shape[1]), int(input_shape[2]), self.n_decoder_channels]
            self.h = rt.add_variable("h", h_shape, rt.rng.uniform(0, 1, size=h_shape))  # decoder.py:60-62

    def __call__(self, prior: DeviceTensor, enc_dec_combiners: List, nll=False, training: bool = False):
        rt = self.rt
        z_params = []
        B = prior.shape[0]
        self.sampler.begin(B)
        # log q / log p are accumulated over groups inside the latent kernel (decoder.py:69-71,84-102)
        log_p = rt.zeros(B)
        log_q = rt.zeros(B)
        lq, lp = (log_q, log_p) if nll else (None, None)
        z0, params = self.sampler(prior, z_idx=0, training=training, log_q=lq, log_p=lp)
        z_params.append(params)
        h = R.broadcast_batch(rt, self.h, B)
        x = self.groups[0](h, z0, training)

        combine_idx = 0
        for group in self.groups[1:]:
            if isinstance(group, DecoderSampleCombiner):
                enc_prior = enc_dec_combiners[combine_idx](x)
                z_sample, params = self.sampler(x, z_idx=combine_idx + 1, enc_prior=enc_prior, training=training,
                                                log_q=lq, log_p=lp)
                z_params.append(params)
                x = group(x, z_sample, training)
                combine_idx += 1
            else:
                x = group(x, training)
        return x, z_params, log_p, log_q

    call = __call__


class DecoderSampleCombiner(Layer):
    def __init__(self, output_channels, *, in_channels: int, name: str = "dsc", **kwargs):
        super().__init__(name)
        with self.rt.scope(name):
            self.conv = SpectralNormalization(
                Conv2D(output_channels, (1, 1), strides=(1, 1), padding="same", in_channels=in_channels))

    def __call__(self, x: DeviceTensor, z: DeviceTensor, training: bool = False) -> DeviceTensor:
        # tf.concat((x, z), axis=3) is a dual-source K loop inside the conv kernel (decoder.py:114-117)
        return self.conv(x, training, x2=z)

    call = __call__


class GenerativeResidualCell(Layer):
    """Generative network residual cell in NVAE architecture (decoder.py:120-147)."""

    def __init__(self, output_channels, expansion_ratio=6, *, name: str = "cell", **kwargs):
        super().__init__(name)
        c, e = output_channels, expansion_ratio * output_channels
        with self.rt.scope(name):
            bn = lambda ch, n: BatchNormalization(momentum=0.05, epsilon=1e-5, channels=ch, name=n, in_bn_loss=True)
            self.batch_norm1 = bn(c, "batch_norm1")
            self.conv1 = SpectralNormalization(Conv2D(e, (1, 1), padding="same", in_channels=c, name="conv1"))
            self.batch_norm2 = bn(e, "batch_norm2")
            self.depth_conv = DepthwiseConv2D((5, 5), padding="same", in_channels=e, name="depth_conv")
            self.batch_norm3 = bn(e, "batch_norm3")
            self.conv2 = SpectralNormalization(Conv2D(c, (1, 1), padding="same", in_channels=e, name="conv2"))
            self.batch_norm4 = bn(c, "batch_norm4")
            self.se = SqueezeExcitation(channels=c, name="se")

    def __call__(self, inputs: DeviceTensor, training: bool = False) -> DeviceTensor:
        rt = self.rt
        # BN1 (no activation) and BN3 + swish are applied inside the 1x1 convolutions that consume them
        x = self.conv1(inputs, training, bn_in=(self.batch_norm1, NVAE_ACT_NONE, training))
        x = R.dwconv_bn_act(rt, x, self.batch_norm2, NVAE_ACT_SWISH, self.depth_conv, training)
        x = self.conv2(x, training, bn_in=(self.batch_norm3, NVAE_ACT_SWISH, training))
        return self.se.fused(x, inputs, 0.1, 1.0, bn=self.batch_norm4, training=training)

    call = __call__
