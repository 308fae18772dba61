#!/usr/bin/env python
"""Convert a checkpoint written by the reference (`model.save_weights(models/epoch_N)`, train.py:13-14,51,57; TensorFlow
object-graph format) into the .npz that `nvae_tf_b200.NVAE.load_weights` reads.

    python tools/convert_tf_checkpoint.py models/epoch_10 epoch_10.npz     # needs TensorFlow at the site that trained

Two parts:
  * `tf_checkpoint_key(arena_name, model)` -- the NAME MAP: arena variable name (this repository's `Runtime.variables`) ->
    key of the same variable in the reference's checkpoint.  The reference builds its blocks from keras.Sequential and
    python lists, so its object-graph paths are `layer_with_weights-<k>` / list-index chains
    (`encoder/groups/3/layer_with_weights-0/conv1/layer/kernel/.ATTRIBUTES/VARIABLE_VALUE`), while the arena uses readable
    names (`encoder/groups/3/cells/0/conv1/kernel`).  The map is pure string logic over the model structure and is
    unit-tested without TensorFlow (tests/test_checkpoint_names.py): it is a bijection onto the arena.
  * `convert(reader, model)` -- pulls every mapped tensor (plus the Adamax slots `.OPTIMIZER_SLOT/optimizer/{m,v}` and
    `optimizer/iter`, when present) out of a `tf.train.load_checkpoint` reader.

Stated plainly: TensorFlow is not installable in this repository's build image, so the key grammar below (TF 2.3 object-graph
naming: attribute names joined by '/', `layer_with_weights-<k>` for the k-th weighted layer of a Sequential, list elements
by index, `tfa.SpectralNormalization` exposing `layer` and `u`) is written from the TensorFlow / TF-Addons sources from memory
and has NOT been run against a real checkpoint here.  `convert` therefore checks every key it expects against the reader's
own variable list and reports all mismatches at once instead of guessing.
"""
from __future__ import annotations

import os
import re
import sys
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
_LEAF = {"kernel": "layer/kernel", "bias": "layer/bias", "u": "u"}  # under a SpectralNormalization wrapper


def _sn(prefix: str, leaf: str) -> str:
    return f"{prefix}/{_LEAF[leaf]}"


def _cell_has_skip(model, tower: str, idx: int) -> bool:
    cell = getattr(model, tower).cells[idx]
    return getattr(cell, "skip", None) is not None


def tf_checkpoint_key(name: str, model) -> str:
    """Arena variable name -> the reference checkpoint's key (without the trailing SUFFIX)."""
    p = name.split("/")
    leaf = p[-1]
    if p[0] == "preprocess":
        root = "preprocess/pre_process"
        if p[1] == "stem":  # Sequential(SpectralNormalization(Conv2D)): the first weighted layer
            return _sn(f"{root}/layer_with_weights-0", leaf)
        cell = f"{root}/layer_with_weights-{int(p[2]) + 1}"  # BNSwishConv cells follow the stem
        if p[3] == "nodes":  # Sequential [BN, swish, SNconv] x n_nodes: weighted layers BN_i = 2i, conv_i = 2i + 1
            i = int(p[4])
            if p[5] == "bn":
                return f"{cell}/nodes/layer_with_weights-{2 * i}/{leaf}"
            return _sn(f"{cell}/nodes/layer_with_weights-{2 * i + 1}", leaf)
        if p[3] == "skip":
            return _sn(f"{cell}/skip/{p[4]}", leaf)
        if p[3] == "se":
            return f"{cell}/se/{p[4]}/{leaf}"
    if p[0] == "encoder":
        if p[1] == "final_enc":  # Sequential [ELU, SNconv, ELU]
            return _sn("encoder/final_enc/layer_with_weights-0", leaf)
        g = f"encoder/groups/{p[2]}"
        if p[3] == "cells":  # a Sequential of EncodingResidualCell
            return _cell_leaf(f"{g}/layer_with_weights-{p[4]}", p[5:])
        if p[3] == "decoder_conv":
            return _sn(f"{g}/decoder_conv", leaf)
        if p[3] == "bn":
            return f"{g}/bn/{leaf}"
        if p[3] == "conv":
            return _sn(f"{g}/conv", leaf)
    if p[0] == "decoder":
        if p[1] == "h":
            return "decoder/h"
        if p[1] == "sampler":
            if p[2] == "enc_sampler":
                return _sn(f"decoder/sampler/enc_sampler/{p[3]}", leaf)
            return _sn(f"decoder/sampler/dec_sampler/{p[3]}/layer_with_weights-0", leaf)  # Sequential [ELU, SNconv]
        g = f"decoder/groups/{p[2]}"
        if p[3] == "cells":
            return _cell_leaf(f"{g}/layer_with_weights-{p[4]}", p[5:])
        if p[3] == "bn":
            return f"{g}/bn/{leaf}"
        if p[3] == "conv":
            return _sn(f"{g}/conv", leaf)
    if p[0] == "postprocess":
        root = "postprocess/sequence"
        if p[1] == "final":  # [cells..., ELU, SNconv]: the weighted layer after the cells
            return _sn(f"{root}/layer_with_weights-{len(model.postprocess.cells)}", leaf)
        k = int(p[2])
        cell = f"{root}/layer_with_weights-{k}"
        if p[3] == "skip":  # Rescaler
            return f"{cell}/skip/bn/{leaf}" if p[4] == "bn" else _sn(f"{cell}/skip/conv", leaf)
        node = f"{cell}/sequence/layer_with_weights-0/sequence"  # PostprocessCell.sequence[0] = PostprocessNode
        up = 1 if _cell_has_skip(model, "postprocess", k) else 0  # an upscaling node starts with a Rescaler
        sub = p[4]
        if sub == "rescaler":
            base = f"{node}/layer_with_weights-0"
            return f"{base}/bn/{leaf}" if p[5] == "bn" else _sn(f"{base}/conv", leaf)
        if sub == "bn0":
            return f"{node}/layer_with_weights-{up}/{leaf}"
        if sub in ("cbs1", "cbs2"):  # ConvBNSwish.sequence = [SNconv, BN, swish]
            base = f"{node}/layer_with_weights-{up + (1 if sub == 'cbs1' else 2)}/sequence"
            return _sn(f"{base}/layer_with_weights-0", leaf) if p[5] == "conv" else f"{base}/layer_with_weights-1/{leaf}"
        if sub == "conv3":
            return _sn(f"{node}/layer_with_weights-{up + 3}", leaf)
        if sub == "bn1":
            return f"{node}/layer_with_weights-{up + 4}/{leaf}"
        if sub == "se":
            return f"{node}/layer_with_weights-{up + 5}/{p[5]}/{leaf}"
    raise KeyError(f"no checkpoint key rule for arena variable {name!r}")


def _cell_leaf(prefix: str, rest: List[str]) -> str:
    """Attributes of EncodingResidualCell / GenerativeResidualCell (encoder.py:91-99, decoder.py:125-136)."""
    attr, leaf = rest[0], rest[-1]
    if attr.startswith("batch_norm"):
        return f"{prefix}/{attr}/{leaf}"
    if attr in ("conv1", "conv2"):
        return _sn(f"{prefix}/{attr}", leaf)
    if attr == "depth_conv":
        return f"{prefix}/depth_conv/{leaf}"
    if attr == "se":
        return f"{prefix}/se/{rest[1]}/{leaf}"
    raise KeyError("/".join([prefix] + rest))


def name_map(model) -> Dict[str, str]:
    """{arena name: full checkpoint key} for every variable of `model` (an nvae_tf_b200.NVAE; device='cpu' is enough)."""
    return {n: "model/" * 0 + tf_checkpoint_key(n, model) + SUFFIX for n in model.rt.variables}


def convert(reader, model) -> Dict[str, np.ndarray]:
    """`reader`: tf.train.load_checkpoint(path) (anything with get_variable_to_shape_map() and get_tensor(key))."""
    have = set(reader.get_variable_to_shape_map())
    out, missing = {}, []
    for name, key in name_map(model).items():
        if key not in have:
            missing.append((name, key))
            continue
        a = np.asarray(reader.get_tensor(key), dtype=np.float32)
        want = model.rt.variables[name].shape
        if tuple(a.shape) != tuple(want):
            raise ValueError(f"{key}: checkpoint shape {a.shape} != model shape {want}")
        out[name] = a
        for slot in ("m", "v"):  # Adamax slots, when the checkpoint carries the optimizer
            sk = key[:-len(SUFFIX)] + f"/.OPTIMIZER_SLOT/optimizer/{slot}" + SUFFIX
            if sk in have:
                out[f"__optimizer/{slot}/{name}"] = np.asarray(reader.get_tensor(sk), dtype=np.float32)
    if missing:
        raise KeyError(f"{len(missing)} expected keys are not in the checkpoint, e.g. {missing[:5]}; checkpoint has e.g. "
                       f"{sorted(have)[:5]}")
    it = "optimizer/iter" + SUFFIX
    if it in have:
        out["__optimizer_iterations"] = np.asarray(reader.get_tensor(it))
    return out


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    try:
        import tensorflow as tf
    except ImportError:
        print("TensorFlow is needed to READ the reference's checkpoint (run this where the model was trained).")
        return 1
    import bench
    from nvae_tf_b200.models import NVAE
    model = NVAE(**bench.mirror_kwargs(1), device="cpu")
    arrays = convert(tf.train.load_checkpoint(argv[1]), model)
    m = re.search(r"epoch_(\d+)", argv[1])
    epoch = int(m.group(1)) if m else 0
    arrays.setdefault("__epoch", np.asarray(epoch))
    arrays.setdefault("__steps", np.asarray(epoch * 417))  # train.py:133-135 re-derives steps the same way
    np.savez(argv[2], **arrays)
    print(f"wrote {len(arrays)} arrays to {argv[2]}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
