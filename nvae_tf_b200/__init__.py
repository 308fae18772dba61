"""nvae_tf_b200: B200 (sm_100a) drop-in for the data-parallel hot path of stevensdavid/nvae-tf.

`models.NVAE`, `common`, `encoder`, `decoder` (+ `preprocess`, `postprocess`) mirror the
reference's Keras classes; every arithmetic op is a libnvae_b200.so launch through the C-ABI of
`include/nvae_b200.h`.  Importing the package does not need a GPU; constructing a model does.
"""
from . import _lib  # noqa: F401
from ._lib import NVAE_PREC_FP32, NVAE_PREC_TF32, NVAE_PREC_TF32X3, NvaeError  # noqa: F401

__all__ = ["NVAE_PREC_FP32", "NVAE_PREC_TF32", "NVAE_PREC_TF32X3", "NvaeError"]
