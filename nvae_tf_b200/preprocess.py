"""Drop-in for the reference's preprocess.py (adjacent row f1 of SURVEY 8f): Preprocess,
SkipScaler, BNSwishConv (preprocess.py:7-107) on the same kernels as the cells."""
from __future__ import annotations

from . import runtime as R
from ._lib import NVAE_ACT_SWISH
from .common import SqueezeExcitation
from .layers import BatchNormalization, Conv2D, Layer, SpectralNormalization
from .runtime import DeviceTensor


class Preprocess(Layer):
    def __init__(self, n_encoder_channels, n_blocks, n_cells, scale_factor, input_shape, mult=1, *,
                 name: str = "preprocess", **kwargs) -> None:
        super().__init__(name)
        input_shape = [int(v) for v in input_shape]
        self.cells = []
        with self.rt.scope(name):
            self.stem = SpectralNormalization(
                Conv2D(n_encoder_channels, (3, 3), padding="same", in_channels=input_shape[3], name="stem"))
            in_channels = n_encoder_channels
            with self.rt.scope("cells"):
                for block in range(n_blocks):
                    for cell in range(n_cells - 1):
                        n_channels = mult * n_encoder_channels
                        self.cells.append(BNSwishConv(2, n_channels, stride=(1, 1), in_channels=in_channels,
                                                      name=str(len(self.cells))))
                        in_channels = n_channels
                    # rescale channels on the final cell of the block (preprocess.py:30-33)
                    n_channels = mult * n_encoder_channels * scale_factor
                    self.cells.append(BNSwishConv(2, n_channels, stride=(2, 2), in_channels=in_channels,
                                                  name=str(len(self.cells))))
                    in_channels = n_channels
                    mult *= scale_factor
                    input_shape = [input_shape[0], input_shape[1] // scale_factor, input_shape[2] // scale_factor,
                                   input_shape[3] * scale_factor]
        self.mult = mult
        self.output_shape_ = input_shape
        self.out_channels = in_channels

    def __call__(self, inputs: DeviceTensor, training: bool = False) -> DeviceTensor:
        # 2 * inputs - 1 ([0,1] -> [-1,1], preprocess.py:39) is folded into the stem conv's operand load
        x = self.stem(inputs, training, pre=(2.0, -1.0))
        for cell in self.cells:
            x = cell(x, training)
        return x

    call = __call__


class SkipScaler(Layer):
    def __init__(self, n_channels, *, in_channels: int, name: str = "skip", **kwargs):
        super().__init__(name)
        q = n_channels // 4
        self.n_channels = n_channels
        with self.rt.scope(name):
            mk = lambda f, n: SpectralNormalization(
                Conv2D(f, (1, 1), strides=(2, 2), padding="same", in_channels=in_channels, name=n))
            # each convolution handles a quarter of the channels; conv4 takes the remainder
            self.conv1, self.conv2, self.conv3 = mk(q, "conv1"), mk(q, "conv2"), mk(q, "conv3")
            self.conv4 = mk(n_channels - 3 * q, "conv4")

    def __call__(self, x: DeviceTensor, training: bool = False) -> DeviceTensor:
        rt = self.rt
        out = R.bn_act(rt, x, None, NVAE_ACT_SWISH, False)
        N, H, W, _ = x.shape
        y = DeviceTensor(rt.empty(N, -(-H // 2), -(-W // 2), self.n_channels))
        q = self.n_channels // 4
        # strided 1x1 convs on shifted views write their channel slice of y directly (no tf.concat copy)
        self.conv1(out, training, out=y, y_off=0)
        self.conv2(out, training, shift=(1, 1), out=y, y_off=q)
        self.conv3(out, training, shift=(0, 1), out=y, y_off=2 * q)
        self.conv4(out, training, shift=(1, 0), out=y, y_off=3 * q)
        return y

    call = __call__


class BNSwishConv(Layer):
    def __init__(self, n_nodes, n_channels, stride, *, in_channels: int, name: str = "cell", **kwargs) -> None:
        super().__init__(name)
        self.bns, self.convs = [], []
        with self.rt.scope(name):
            if stride == (1, 1):
                self.skip = None  # tf.identity
            elif stride == (2, 2):
                self.skip = SkipScaler(n_channels, in_channels=in_channels, name="skip")
            cin = in_channels
            with self.rt.scope("nodes"):
                for i in range(n_nodes):
                    with self.rt.scope(str(i)):
                        self.bns.append(BatchNormalization(momentum=0.05, epsilon=1e-5, channels=cin, name="bn"))
                        self.convs.append(SpectralNormalization(
                            Conv2D(n_channels, (3, 3), stride if i == 0 else (1, 1), padding="same",
                                   in_channels=cin, name="conv")))
                    cin = n_channels
            self.se = SqueezeExcitation(channels=n_channels, name="se")

    def __call__(self, inputs: DeviceTensor, training: bool = False) -> DeviceTensor:
        rt = self.rt
        skipped = inputs if self.skip is None else self.skip(inputs, training)
        x = inputs
        for bn, conv in zip(self.bns, self.convs):
            x = R.bn_act(rt, x, bn, NVAE_ACT_SWISH, training)
            x = conv(x, training)
        return self.se.fused(x, skipped, 1.0, 0.1, training=training)  # skipped + 0.1 * se(x)

    call = __call__
