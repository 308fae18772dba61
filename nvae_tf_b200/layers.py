"""Stand-ins for the `tensorflow.keras.layers` / `tensorflow_addons` objects the reference builds
its blocks from (common.py:7-8, encoder.py:3-6, decoder.py:3-5): same constructor arguments and
variable names/shapes (kernel HWIO, depthwise_kernel [5,5,C,1], Dense kernel [in,out], BN
gamma/beta/moving_mean/moving_variance, SN `u` [1,Cout]) so reference checkpoints map 1:1.
Unlike Keras these are built eagerly (`in_channels` is explicit) because all variables must be
laid out in the flat arenas before the first step.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import runtime as R
from ._lib import NVAE_ACT_ELU, NVAE_ACT_NONE, NVAE_ACT_SWISH
from .runtime import DeviceTensor


def _pair(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


class Layer:
    def __init__(self, name: Optional[str] = None):
        self.rt = R.current()
        self.name = name

    def scope(self, name: Optional[str] = None):
        return self.rt.scope(name if name is not None else (self.name or ""))


class Conv2D(Layer):
    """layers.Conv2D(filters, kernel_size, strides, padding='same', use_bias) -- NHWC x HWIO."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", use_bias=True, *, in_channels: int,
                 name: str = "conv"):
        super().__init__(name)
        self.filters = int(filters)  # Keras casts float channel counts (decoder.py:44,52)
        kh, kw = _pair(kernel_size)
        self.stride = _pair(strides)[0]
        if padding == "valid" and (kh, kw) != (1, 1):
            raise ValueError("only 1x1 convolutions may use padding='valid' on this path")
        rt = self.rt
        with rt.scope(name):
            fan_in, fan_out = kh * kw * in_channels, kh * kw * self.filters
            self.kernel = rt.add_variable("kernel", (kh, kw, in_channels, self.filters),
                                          rt.glorot((kh, kw, in_channels, self.filters), fan_in, fan_out))
            self.bias = rt.add_variable("bias", (self.filters,), np.zeros(self.filters)) if use_bias else None
        # TF32 operand copies of the kernel inside rt.pack (tensor-core modes only; laid out by Runtime.finalize)
        self.rnd_off = self.tr_off = -1
        self.sn = None  # the SpectralNormalization wrapper, if any
        rt.convs.append(self)

    def packed_fwd(self):
        """[Cout][tap][Cin] TF32 copy: the K-major B operand of the forward GEMM (None in fp32 mode)."""
        return self.rt.pack[self.tr_off:].data_ptr() if self.tr_off >= 0 else None

    def packed_dgrad(self):
        """HWIO operand of the backward-data GEMM (K-major along Cout): the TF32-rounded copy in NVAE_PREC_TF32,
        the master kernel itself in NVAE_PREC_TF32X3, None in fp32 mode."""
        if self.rnd_off >= 0:
            return self.rt.pack[self.rnd_off:].data_ptr()
        return self.kernel.ptr() if self.tr_off >= 0 else None

    def __call__(self, x: DeviceTensor, x2=None, residual=None, bn_in=None, **kw) -> DeviceTensor:
        """bn_in=(bn, act, training): the input is act(bn(x)), applied inside the convolution where the kernel can."""
        if self.sn is None and self.tr_off >= 0 and not self.rt.sn_done:
            self.rt.pack_plain(self)  # un-wrapped conv (tests): refresh the operand copies on every call
        if bn_in is not None:
            if x2 is not None or kw:
                raise NotImplementedError("bn_in with a second source / shifted view")
            return R.bn_conv2d(self.rt, x, bn_in[0], bn_in[1], self, bn_in[2], residual=residual)
        return R.conv2d(self.rt, x, self, x2=x2, residual=residual, **kw)


class SpectralNormalization(Layer):
    """tfa.layers.SpectralNormalization(layer, power_iterations=1)."""

    def __init__(self, layer: Conv2D):
        super().__init__(layer.name)
        self.layer = layer
        rt = self.rt
        with rt.scope(layer.name):
            u = np.clip(rt.rng.normal(0, 0.02, size=(1, layer.filters)), -0.04, 0.04)  # TruncatedNormal(0.02)
            self.u = rt.add_variable("u", (1, layer.filters), u, trainable=False)
        self.index = -1
        layer.sn = self
        rt.sn_convs.append(self)

    def __call__(self, x: DeviceTensor, training: bool = False, **kw) -> DeviceTensor:
        if not self.rt.sn_done and (training or self.layer.tr_off >= 0):
            self.rt.spectral_normalize_one(self.index, power_iter=training)
        return self.layer(x, **kw)


class DepthwiseConv2D(Layer):
    """layers.DepthwiseConv2D((5,5), padding='same') (decoder.py:130)."""

    def __init__(self, kernel_size=(5, 5), padding="same", *, in_channels: int, name: str = "depth_conv"):
        super().__init__(name)
        if _pair(kernel_size) != (5, 5) or padding != "same":
            raise ValueError("only the 5x5 'same' depthwise convolution of decoder.py:130 is on this path")
        rt = self.rt
        with rt.scope(name):
            self.depthwise_kernel = rt.add_variable("depthwise_kernel", (5, 5, in_channels, 1),
                                                    rt.glorot((5, 5, in_channels, 1), 25 * in_channels, 25))
            self.bias = rt.add_variable("bias", (in_channels,), np.zeros(in_channels))


class BatchNormalization(Layer):
    """layers.BatchNormalization(momentum=0.05, epsilon=1e-5), axis=-1."""

    def __init__(self, momentum=0.99, epsilon=1e-3, *, channels: int, name: str = "bn", in_bn_loss: bool = False):
        super().__init__(name)
        self.momentum, self.epsilon = float(momentum), float(epsilon)
        rt = self.rt
        with rt.scope(name):
            self.gamma = rt.add_variable("gamma", (channels,), np.ones(channels))
            self.beta = rt.add_variable("beta", (channels,), np.zeros(channels))
            self.moving_mean = rt.add_variable("moving_mean", (channels,), np.zeros(channels), trainable=False)
            self.moving_variance = rt.add_variable("moving_variance", (channels,), np.ones(channels), trainable=False)
        if in_bn_loss:
            rt.bn_loss_layers.append(self)

    @property
    def weights(self):  # models.py:258 reads layer.weights[0] == gamma
        return [self.gamma, self.beta, self.moving_mean, self.moving_variance]

    def __call__(self, x: DeviceTensor, training: bool = False) -> DeviceTensor:
        return R.bn_act(self.rt, x, self, NVAE_ACT_NONE, training)


class Dense(Layer):
    """layers.Dense(units): variables only -- the SE kernels consume them (common.py:126-127)."""

    def __init__(self, units, *, in_features: int, name: str):
        super().__init__(name)
        self.units = int(units)
        rt = self.rt
        with rt.scope(name):
            self.kernel = rt.add_variable("kernel", (in_features, self.units),
                                          rt.glorot((in_features, self.units), in_features, self.units))
            self.bias = rt.add_variable("bias", (self.units,), np.zeros(self.units))


class activations:
    """tensorflow.keras.activations.{swish, elu} as bare (un-normalised) activation launches."""

    @staticmethod
    def swish(x: DeviceTensor) -> DeviceTensor:
        return R.bn_act(R.current(), x, None, NVAE_ACT_SWISH, False)

    @staticmethod
    def elu(x: DeviceTensor) -> DeviceTensor:
        return R.bn_act(R.current(), x, None, NVAE_ACT_ELU, False)


class ELU(Layer):
    def __call__(self, x: DeviceTensor, training: bool = False) -> DeviceTensor:
        return R.bn_act(self.rt, x, None, NVAE_ACT_ELU, False)
