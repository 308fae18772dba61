"""Drop-in for the reference's postprocess.py (adjacent row f1 of SURVEY 8f): Postprocess,
PostprocessCell, PostprocessNode, ConvBNSwish (postprocess.py:8-111) on the same kernels."""
from __future__ import annotations

from . import runtime as R
from ._lib import NVAE_ACT_ELU, NVAE_ACT_NONE, NVAE_ACT_SWISH
from .common import RescaleType, Rescaler, SqueezeExcitation
from .layers import BatchNormalization, Conv2D, Layer, SpectralNormalization
from .runtime import DeviceTensor


class Postprocess(Layer):
    def __init__(self, n_blocks, n_cells, mult, n_channels_decoder, scale_factor, *, in_channels: int,
                 out_channels: int = 1, name: str = "postprocess", **kwargs) -> None:
        super().__init__(name)
        self.cells = []
        with self.rt.scope(name):
            with self.rt.scope("cells"):
                for block in range(n_blocks):
                    mult /= scale_factor  # the first cell of each block rescales
                    output_channels = int(n_channels_decoder * mult)
                    for cell_idx in range(n_cells):
                        self.cells.append(PostprocessCell(output_channels, n_nodes=1, upscale=cell_idx == 0,
                                                          scale_factor=scale_factor, in_channels=in_channels,
                                                          name=str(len(self.cells))))
                        in_channels = output_channels
            self.final = SpectralNormalization(
                Conv2D(out_channels, kernel_size=(3, 3), padding="same", in_channels=in_channels, name="final"))
        self.mult = mult

    def __call__(self, inputs: DeviceTensor, training: bool = False) -> DeviceTensor:
        x = inputs
        for cell in self.cells:
            x = cell(x, training)
        x = R.bn_act(self.rt, x, None, NVAE_ACT_ELU, False)
        return self.final(x, training)

    call = __call__


class PostprocessCell(Layer):
    def __init__(self, n_channels, n_nodes, scale_factor, upscale=False, *, in_channels: int, name: str = "cell",
                 **kwargs) -> None:
        super().__init__(name)
        if n_nodes != 1:
            raise ValueError("the reference only ever builds n_nodes=1 (postprocess.py:21)")
        with self.rt.scope(name):
            self.skip = Rescaler(n_channels, scale_factor=scale_factor, rescale_type=RescaleType.UP,
                                 in_channels=in_channels, name="skip") if upscale else None
            self.node = PostprocessNode(n_channels, upscale=upscale, scale_factor=scale_factor,
                                        in_channels=in_channels, name="node")

    def __call__(self, inputs: DeviceTensor, training: bool = False) -> DeviceTensor:
        skipped = inputs if self.skip is None else self.skip(inputs, training)
        return self.node(inputs, skipped, training)  # skip(inputs) + 0.1 * sequence(inputs)

    call = __call__


class ConvBNSwish(Layer):
    """SN(Conv2D(use_bias=False)) -> BN -> swish.  The BN+swish is applied by the consumer of the
    returned pair (fused into the next operand staging), so __call__ returns (conv_out, bn)."""

    def __init__(self, n_channels, kernel_size, stride, groups=1, *, in_channels: int, name: str, **kwargs) -> None:
        super().__init__(name)
        with self.rt.scope(name):
            self.conv = SpectralNormalization(
                Conv2D(n_channels, kernel_size=kernel_size, strides=stride, use_bias=False, padding="same",
                       in_channels=in_channels, name="conv"))
            self.bn = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=n_channels, name="bn")

    def __call__(self, x: DeviceTensor, training: bool = False, bn_in=None, defer: bool = False) -> DeviceTensor:
        """bn_in: a BatchNorm (+ activation) in front of the conv, applied inside it; defer: return the raw conv output --
        the caller hands (self.bn, swish) to the next 1x1 convolution as ITS bn_in."""
        x = self.conv(x, training, bn_in=bn_in) if bn_in is not None else self.conv(x, training)
        return x if defer else R.bn_act(self.rt, x, self.bn, NVAE_ACT_SWISH, training)

    call = __call__


class PostprocessNode(Layer):
    def __init__(self, n_channels, scale_factor, upscale=False, expansion_ratio=6, *, in_channels: int,
                 name: str = "node", **kwargs) -> None:
        super().__init__(name)
        hidden_dim = n_channels * expansion_ratio
        with self.rt.scope(name):
            self.rescaler = Rescaler(n_channels, scale_factor, rescale_type=RescaleType.UP, in_channels=in_channels,
                                     name="rescaler") if upscale else None
            self.bn0 = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=n_channels, name="bn0")
            self.cbs1 = ConvBNSwish(hidden_dim, kernel_size=(1, 1), stride=(1, 1), in_channels=n_channels, name="cbs1")
            # the reference commented `groups=` out, so this is a FULL 5x5 conv (postprocess.py:74-76)
            self.cbs2 = ConvBNSwish(hidden_dim, kernel_size=(5, 5), stride=(1, 1), in_channels=hidden_dim, name="cbs2")
            self.conv3 = SpectralNormalization(
                Conv2D(n_channels, kernel_size=(1, 1), strides=(1, 1), use_bias=False, in_channels=hidden_dim,
                       name="conv3"))
            self.bn1 = BatchNormalization(momentum=0.05, epsilon=1e-5, channels=n_channels, name="bn1")
            self.se = SqueezeExcitation(channels=n_channels, name="se")

    def __call__(self, inputs: DeviceTensor, skipped: DeviceTensor, training: bool = False) -> DeviceTensor:
        rt = self.rt
        x = inputs if self.rescaler is None else self.rescaler(inputs, training)
        x = self.cbs1(x, training, bn_in=(self.bn0, NVAE_ACT_NONE, training))
        x = self.cbs2(x, training, defer=True)
        x = self.conv3(x, training, bn_in=(self.cbs2.bn, NVAE_ACT_SWISH, training))
        return self.se.fused(x, skipped, 1.0, 0.1, bn=self.bn1, training=training)

    call = __call__
