"""Drop-in for the reference's encoder.py: EncoderDecoderCombiner, Encoder, EncodingResidualCell
(encoder.py:9-107) on libnvae_b200 kernels."""
from __future__ import annotations

from functools import partial
from typing import List

from . import runtime as R
from ._lib import NVAE_ACT_ELU, NVAE_ACT_SWISH
from .common import RescaleType, Rescaler, SqueezeExcitation
from .layers import BatchNormalization, Conv2D, Layer, SpectralNormalization
from .runtime import DeviceTensor


class EncoderDecoderCombiner(Layer):
    def __init__(self, n_channels, *, name: str = "combiner", **kwargs) -> None:
        super().__init__(name)
        with self.rt.scope(name):
            self.decoder_conv = SpectralNormalization(
                Conv2D(n_channels, (1, 1), in_channels=n_channels, name="decoder_conv"))

    def __call__(self, encoder_x: DeviceTensor, decoder_x: DeviceTensor, training: bool = False) -> DeviceTensor:
        # encoder_x + conv1x1(decoder_x): the add is the conv kernel's residual epilogue (encoder.py:14-16)
        return self.decoder_conv(decoder_x, training, residual=encoder_x)

    call = __call__


class _Sequential(Layer):
    """keras.Sequential of cells; `.layers` is what calculate_bn_loss recurses into (models.py:259-261)."""

    def __init__(self, name: str):
        super().__init__(name)
        self.layers: List[Layer] = []

    def add(self, layer):
        self.layers.append(layer)

    def __call__(self, x, training: bool = False):
        for layer in self.layers:
            x = layer(x, training)
        return x


class Encoder(Layer):
    def __init__(self, n_encoder_channels, n_latent_per_group: int, res_cells_per_group, n_latent_scales: int,
                 n_groups_per_scale: List[int], mult: int, scale_factor: int, input_shape, *, name: str = "encoder",
                 **kwargs):
        super().__init__(name)
        self.groups = []
        input_shape = list(input_shape)
        with self.rt.scope(name), self.rt.scope("groups"):
            for scale in range(n_latent_scales):
                n_groups = n_groups_per_scale[scale]
                for group_idx in range(n_groups):
                    output_channels = n_encoder_channels * mult
                    group = _Sequential(str(len(self.groups)))
                    with self.rt.scope(group.name), self.rt.scope("cells"):
                        for k in range(res_cells_per_group):
                            group.add(EncodingResidualCell(output_channels, name=str(k)))
                    self.groups.append(group)
                    if not (scale == n_latent_scales - 1 and group_idx == n_groups - 1):
                        # a convolution between each group except the final output (encoder.py:44-46)
                        self.groups.append(EncoderDecoderCombiner(output_channels, name=str(len(self.groups))))
                if scale < n_latent_scales - 1:  # downsample at the end of each scale except the last
                    output_channels = n_encoder_channels * mult * scale_factor
                    self.groups.append(Rescaler(output_channels, scale_factor=scale_factor,
                                                rescale_type=RescaleType.DOWN, in_channels=n_encoder_channels * mult,
                                                name=str(len(self.groups)), in_bn_loss=True))
                    mult *= scale_factor
                    input_shape = [input_shape[0], input_shape[1] // scale_factor, input_shape[2] // scale_factor,
                                   input_shape[3] * scale_factor]
        with self.rt.scope(name), self.rt.scope("final_enc"):
            # Sequential([ELU, SN(Conv2D(C,(1,1))), ELU]) encoder.py:58-66
            self.final_enc_conv = SpectralNormalization(
                Conv2D(n_encoder_channels * mult, (1, 1), padding="same", in_channels=n_encoder_channels * mult))
        self.mult = mult
        self.output_shape_ = input_shape

    def final_enc(self, x: DeviceTensor, training: bool = False) -> DeviceTensor:
        x = R.bn_act(self.rt, x, None, NVAE_ACT_ELU, False)
        x = self.final_enc_conv(x, training)
        return R.bn_act(self.rt, x, None, NVAE_ACT_ELU, False)

    def __call__(self, x: DeviceTensor, training: bool = False):
        enc_dec_combiners = []
        for group in self.groups:
            if isinstance(group, EncoderDecoderCombiner):
                # stepping between groups: defer the combiner with the encoder activation bound (encoder.py:74-79)
                enc_dec_combiners.append(partial(group, x, training=training))
            else:
                x = group(x, training)
        final = self.final_enc(x, training)
        return enc_dec_combiners, final

    call = __call__


class EncodingResidualCell(Layer):
    """Encoding network residual cell in NVAE architecture (encoder.py:86-107)."""

    def __init__(self, output_channels, *, name: str = "cell", **kwargs):
        super().__init__(name)
        c = output_channels
        with self.rt.scope(name):
            self.batch_norm1 = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=c, name="batch_norm1",
                                                  in_bn_loss=True)
            self.conv1 = SpectralNormalization(Conv2D(c, (3, 3), padding="same", in_channels=c, name="conv1"))
            self.batch_norm2 = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=c, name="batch_norm2",
                                                  in_bn_loss=True)
            self.conv2 = SpectralNormalization(Conv2D(c, (3, 3), padding="same", in_channels=c, name="conv2"))
            self.se = SqueezeExcitation(channels=c, name="se")

    def __call__(self, inputs: DeviceTensor, training: bool = False) -> DeviceTensor:
        rt = self.rt
        x = R.bn_act(rt, inputs, self.batch_norm1, NVAE_ACT_SWISH, training)
        x = self.conv1(x, training)
        x = R.bn_act(rt, x, self.batch_norm2, NVAE_ACT_SWISH, training)
        x = self.conv2(x, training)
        return self.se.fused(x, inputs, 0.1, 1.0, training=training)  # 0.1 * inputs + se(x)

    call = __call__
