"""Keras numbers auto-named layers per session, so a decoder saved from the notebook that also built the encoder has
'dense_1', 'conv2d_transpose_5', ...: the weight loader must find layers by kernel shape, not by name."""
import os

import numpy as np
import pytest

from srcfd import sr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_decoder_layers_resolved_by_shape_under_session_numbering():
    w = sr.glorot_decoder_weights(0)
    m = {"dense": "dense_1", "conv2d_transpose": "conv2d_transpose_5", "conv2d_transpose_1": "conv2d_transpose_6",
         "conv2d_transpose_2": "conv2d_transpose_7", "conv2d_transpose_3": "conv2d_transpose_8",
         "conv2d_transpose_4": "conv2d_transpose_9", "output_image_400": "output_image_400"}
    ren = {f"{m[k.split('/')[0]]}/{k.split('/')[1]}": v for k, v in w.items()}
    out = sr.resolve_layers(ren, sr.DECODER_LAYERS, sr.DECODER_SHAPES)
    assert set(out) == set(w) and all(np.array_equal(out[k], w[k]) for k in w)
    bad = dict(ren); bad.pop("conv2d_transpose_7/kernel")
    with pytest.raises(ValueError, match=r"no layer with a \(2, 2, 32, 64\) kernel"):
        sr.resolve_layers(bad, sr.DECODER_LAYERS, sr.DECODER_SHAPES)
    wrong = dict(w); wrong["dense/kernel"] = np.zeros((50, 7), np.float32)
    with pytest.raises(ValueError, match="dense/kernel has shape"):
        sr.resolve_layers(wrong, sr.DECODER_LAYERS, sr.DECODER_SHAPES)


def test_committed_encoder_file_resolves():
    enc = sr.read_keras_weights(os.path.join(GOLD, "encoder10_multiBC.h5"))
    out = sr.resolve_layers(enc, sr.ENCODER_LAYERS, sr.ENCODER_SHAPES)
    for name in sr.ENCODER_LAYERS:
        assert out[f"{name}/kernel"].shape == sr.ENCODER_SHAPES[name]
