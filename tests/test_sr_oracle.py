"""SR autoencoder restatement (oracle/sr_oracle.py) against an independent torch-CPU implementation,
plus the weight reader on the real committed encoder file.  CPU only."""
import os

import numpy as np
import pytest

from oracle import sr_oracle as S
from srcfd import sr


def test_encoder_real_weights_numpy_vs_torch(golden_dir):
    w = sr.read_keras_weights(os.path.join(golden_dir, "encoder10_multiBC.h5"))
    assert w["conv2d/kernel"].shape == (3, 3, 1, 64) and w["latent_vector/kernel"].shape == (128, 50)
    assert sum(v.size for v in w.values()) == 490674
    x = np.random.default_rng(0).standard_normal((5, 10, 10, 1)).astype(np.float32)
    a, b = S.encoder_forward(x, w), S.torch_encoder_forward(x, w)
    assert a.shape == (5, 50)
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-5)


def test_decoder_synthetic_numpy_vs_torch():
    w = sr.glorot_decoder_weights(0)
    assert sum(v.size for v in w.values()) == 2218817
    z = np.random.default_rng(1).standard_normal((2, 50)).astype(np.float32)
    a, b = S.decoder_forward(z, w), S.torch_decoder_forward(z, w)
    assert a.shape == (2, 400, 400, 1)
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=2e-5)


def test_transposed_conv_semantics_small():
    """out[y*s+ky, x*s+kx, co] += in[y,x,ci] * W[ky,kx,co,ci] checked by brute force."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((1, 3, 4, 2)).astype(np.float32); W = rng.standard_normal((3, 3, 5, 2)).astype(np.float32)
    b = rng.standard_normal(5).astype(np.float32)
    ref = np.zeros((1, 7, 9, 5), dtype=np.float64)
    for y in range(3):
        for xx in range(4):
            for ky in range(3):
                for kx in range(3):
                    ref[0, y * 2 + ky, xx * 2 + kx] += W[ky, kx].astype(np.float64) @ x[0, y, xx].astype(np.float64)
    np.testing.assert_allclose(S.conv2d_transpose_valid(x, W, b, 2), ref + b, rtol=1e-5, atol=1e-5)
