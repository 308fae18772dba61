"""Name map of tools/convert_tf_checkpoint.py (arena variable -> key in the reference's TF object-graph checkpoint,
train.py:13-14,51): checked on names only -- no TensorFlow here."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import convert_tf_checkpoint as K  # noqa: E402


@pytest.fixture(scope="module")
def model():
    import bench
    from nvae_tf_b200.models import NVAE
    return NVAE(**bench.mirror_kwargs(2), device="cpu")


def test_every_arena_variable_has_a_distinct_checkpoint_key(model):
    nm = K.name_map(model)
    assert len(nm) == len(model.rt.variables) and len(set(nm.values())) == len(nm)  # a bijection onto the arena
    assert all(k.endswith(K.SUFFIX) for k in nm.values())
    for key in nm.values():  # only attribute names of the reference's classes, list indices and Sequential slots
        for part in key[:-len(K.SUFFIX)].split("/"):
            assert part.isdigit() or part.startswith("layer_with_weights-") or part in {
                "preprocess", "pre_process", "nodes", "skip", "se", "dense1", "dense2", "conv1", "conv2", "conv3", "conv4",
                "encoder", "groups", "final_enc", "decoder_conv", "bn", "conv", "decoder", "sampler", "enc_sampler",
                "dec_sampler", "h", "postprocess", "sequence", "batch_norm1", "batch_norm2", "batch_norm3", "batch_norm4",
                "depth_conv", "layer", "kernel", "bias", "u", "gamma", "beta", "moving_mean", "moving_variance",
                "depthwise_kernel"}, key


def test_known_keys(model):
    k = lambda n: K.tf_checkpoint_key(n, model)
    # encoder.py:39-45: groups = [Sequential(cell), combiner, Sequential(cell), ...]; SN wrapper exposes layer / u
    assert k("encoder/groups/0/cells/0/conv1/kernel") == "encoder/groups/0/layer_with_weights-0/conv1/layer/kernel"
    assert k("encoder/groups/0/cells/0/conv1/u") == "encoder/groups/0/layer_with_weights-0/conv1/u"
    assert k("encoder/groups/1/decoder_conv/bias") == "encoder/groups/1/decoder_conv/layer/bias"
    assert k("encoder/groups/0/cells/0/se/dense2/kernel") == "encoder/groups/0/layer_with_weights-0/se/dense2/kernel"
    assert k("encoder/final_enc/conv/kernel") == "encoder/final_enc/layer_with_weights-0/layer/kernel"
    # decoder.py:37-41,60-62; common.py:39-63
    assert k("decoder/groups/1/cells/0/depth_conv/depthwise_kernel") == \
        "decoder/groups/1/layer_with_weights-0/depth_conv/depthwise_kernel"
    assert k("decoder/sampler/dec_sampler/3/conv/kernel") == "decoder/sampler/dec_sampler/3/layer_with_weights-0/layer/kernel"
    assert k("decoder/sampler/enc_sampler/0/u") == "decoder/sampler/enc_sampler/0/u"
    assert k("decoder/h") == "decoder/h"
    # preprocess.py:19-33,80-101: pre_process = Sequential(stem, cells...); nodes = [BN, swish, SNconv] x 2
    assert k("preprocess/stem/kernel") == "preprocess/pre_process/layer_with_weights-0/layer/kernel"
    assert k("preprocess/cells/0/nodes/1/bn/gamma") == "preprocess/pre_process/layer_with_weights-1/nodes/layer_with_weights-2/gamma"
    assert k("preprocess/cells/2/nodes/0/conv/kernel") == \
        "preprocess/pre_process/layer_with_weights-3/nodes/layer_with_weights-1/layer/kernel"
    assert k("preprocess/cells/2/skip/conv4/u") == "preprocess/pre_process/layer_with_weights-3/skip/conv4/u"
    # postprocess.py:13-29,42-54,66-88,94-108: upscaling node = [Rescaler, BN, cbs1, cbs2, SNconv, BN, SE]
    up = "postprocess/sequence/layer_with_weights-0/sequence/layer_with_weights-0/sequence"
    assert k("postprocess/cells/0/node/rescaler/conv/kernel") == up + "/layer_with_weights-0/conv/layer/kernel"
    assert k("postprocess/cells/0/node/bn0/gamma") == up + "/layer_with_weights-1/gamma"
    assert k("postprocess/cells/0/node/cbs2/conv/kernel") == up + "/layer_with_weights-3/sequence/layer_with_weights-0/layer/kernel"
    assert k("postprocess/cells/0/node/cbs2/bn/beta") == up + "/layer_with_weights-3/sequence/layer_with_weights-1/beta"
    assert k("postprocess/cells/0/node/se/dense1/bias") == up + "/layer_with_weights-6/dense1/bias"
    assert k("postprocess/cells/0/skip/bn/moving_mean") == "postprocess/sequence/layer_with_weights-0/skip/bn/moving_mean"
    flat = "postprocess/sequence/layer_with_weights-1/sequence/layer_with_weights-0/sequence"
    assert k("postprocess/cells/1/node/bn0/gamma") == flat + "/layer_with_weights-0/gamma"
    assert k("postprocess/cells/1/node/conv3/kernel") == flat + "/layer_with_weights-3/layer/kernel"
    assert k("postprocess/final/kernel") == "postprocess/sequence/layer_with_weights-6/layer/kernel"


def test_convert_with_a_stand_in_reader_round_trips(model, tmp_path):
    """`convert` against a fake reader holding exactly the mapped keys (+ Adamax slots for one variable)."""
    nm = K.name_map(model)
    rng = np.random.default_rng(0)
    store = {key: rng.normal(size=model.rt.variables[n].shape).astype(np.float32) for n, key in nm.items()}
    one = "decoder/h"
    base = nm[one][:-len(K.SUFFIX)]
    store[base + "/.OPTIMIZER_SLOT/optimizer/m" + K.SUFFIX] = np.ones(model.rt.variables[one].shape, np.float32)
    store["optimizer/iter" + K.SUFFIX] = np.asarray(1234)

    class Reader:
        def get_variable_to_shape_map(self):
            return {k: v.shape for k, v in store.items()}

        def get_tensor(self, key):
            return store[key]
    out = K.convert(Reader(), model)
    assert set(model.rt.variables) <= set(out) and int(out["__optimizer_iterations"]) == 1234
    assert np.array_equal(out["encoder/groups/3/decoder_conv/kernel"], store[nm["encoder/groups/3/decoder_conv/kernel"]])
    assert out["__optimizer/m/" + one].shape == model.rt.variables[one].shape
    del store[nm["decoder/h"]]
    with pytest.raises(KeyError):
        K.convert(Reader(), model)
