"""CUDA SR autoencoder (through the C ABI) against the CPU restatement.  `-m gpu`.
fp32 results throughout (default precision: split-operand tensor cores, 'bf16x3'; the fp32 CUDA-core path is held to the
same bar); tolerance 1e-4 relative / 5e-5 absolute (different summation order, expf vs np.exp)."""
import os

import numpy as np
import pytest

from oracle import sr_oracle as S

pytestmark = pytest.mark.gpu


def test_encoder_real_weights(golden_dir):
    from srcfd import sr
    enc = sr.load_model(os.path.join(golden_dir, "encoder10_multiBC.h5"))
    x = np.random.default_rng(0).standard_normal((7, 10, 10, 1)).astype(np.float32)
    np.testing.assert_allclose(enc.predict(x), S.encoder_forward(x, enc.weights), rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("B", [1, 3, 40])
def test_decoder_synthetic(B):
    from srcfd import sr
    dec = sr.synthetic_decoder(0)
    z = np.random.default_rng(B).standard_normal((B, 50)).astype(np.float32)
    out = dec.predict(z)
    assert out.shape == (B, 400, 400, 1) and out.dtype == np.float32
    nref = min(B, 3)
    np.testing.assert_allclose(out[:nref], S.decoder_forward(z[:nref], dec.weights), rtol=1e-4, atol=5e-5)
    if B > 3:   # batch independence for the samples not checked against the oracle
        np.testing.assert_array_equal(out[5], dec.predict(z[5:6])[0])


def test_super_resolution_workflow(golden_dir):
    """ml_super_resolution (bfs_ml_accelerated.py:979-1137) end to end vs the same steps on the CPU restatement."""
    from srcfd import bfs, sr
    rng = np.random.default_rng(3)
    coarse = {c: 0.2 * rng.standard_normal((10, 10)) for c in "uvp"}
    stats = os.path.join(golden_dir, "stats_10to400_multiBC.txt")
    encf = os.path.join(golden_dir, "encoder10_multiBC.h5")
    dec = sr.synthetic_decoder(0)
    bfs._wf.verbose = False
    hr = bfs.ml_super_resolution(coarse, 10, 400, stats, encf, dec, use_aspect_ratio_correction=True, lx=10.0, ly=3.0)
    assert set(hr) == set("uvp") and hr["u"].shape == (400, 400)
    # CPU restatement of the same pipeline
    from srcfd.workflow import load_stats, reshape_rectangular_to_square, reshape_square_to_rectangular, standardize_with_stats, inverse_standardize
    lr, hrs = load_stats(stats, 10, 400)
    sq = reshape_rectangular_to_square(coarse, 10, 10, 10.0, 3.0)
    enc_w = sr.read_keras_weights(encf)
    ref = {}
    for c in "uvp":
        x = sq[c].astype(np.float32)
        m = 0.7 * lr[c][0] + 0.3 * np.mean(x); s = 0.7 * lr[c][1] + 0.3 * max(np.std(x), 1e-8)
        xn = standardize_with_stats(x, m, s)[None, ..., None]
        y = S.decoder_forward(S.encoder_forward(xn, enc_w), dec.weights)[0, ..., 0]
        ref[c] = inverse_standardize(y, hrs[c][0], hrs[c][1])
    ref = reshape_square_to_rectangular(ref, 400, 400, 10.0, 3.0)
    for c in "uvp":
        np.testing.assert_allclose(hr[c], ref[c], rtol=2e-4, atol=1e-4)


def _bf16_round(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("layer", [1, 2, 3, 4])
def test_tensor_core_convT_layer(layer):
    """tcgen05 implicit-GEMM ConvT vs the CPU restatement with the SAME bf16-rounded operands: only the fp32
    accumulation order and the final bf16 rounding of the output differ (tolerance: 1 bf16 ulp ~ 0.8 % relative)."""
    from srcfd import sr
    dec = sr.synthetic_decoder(0)
    name = sr.DECODER_LAYERS[layer + 1]
    cin = sr.DECODER_SHAPES[name][3]
    H = 25 * 2 ** (layer - 1)
    B = 2 if layer < 4 else 1
    x = np.random.default_rng(layer).standard_normal((B, H, H, cin)).astype(np.float32)
    got = sr.debug_convT_tc(dec, layer, x)
    assert not sr.tc_error()
    ref = S.swish(S.conv2d_transpose_valid(_bf16_round(x), _bf16_round(dec.weights[f"{name}/kernel"]), dec.weights[f"{name}/bias"], 2))
    np.testing.assert_allclose(got, ref, rtol=1e-2, atol=1e-2)
    assert np.abs(got - ref).max() < 0.02 and np.abs(ref).max() > 0.1


def test_decoder_bf16_tensor_core_path():
    """Whole decoder with the tensor-core layers: bf16 operands => compare with the fp32 restatement at 3e-2 of the
    output range (the throughput option; the default is the split-operand path, held to the fp32 bar)."""
    from srcfd import sr
    dec = sr.synthetic_decoder(0)
    z = np.random.default_rng(11).standard_normal((3, 50)).astype(np.float32)
    sr.set_precision("bf16")
    try:
        out = dec.predict(z)
    finally:
        sr.set_precision("bf16x3")                         # back to the library's default
    assert not sr.tc_error()
    ref = S.decoder_forward(z, dec.weights)
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() <= 3e-2 * scale, (np.abs(out - ref).max(), scale)


def test_final_conv_on_tensor_cores_matches_cuda_core_kernel():
    """bf16 path: the tcgen05 implicit-GEMM final 3x3 conv against the CUDA-core tile kernel on the same bf16
    activation.  The only difference is the bf16 rounding of the 72 weights (the CUDA-core kernel keeps them fp32):
    <= 2^-8 relative per product, bounded here by sum |w| * max |activation| * 2^-8."""
    import os
    from srcfd import sr
    dec = sr.synthetic_decoder(0)
    z = np.random.default_rng(5).standard_normal((2, 50)).astype(np.float32)
    try:
        sr.set_precision("bf16_cc_final")
        ref = dec.predict(z)
        sr.set_precision("bf16")
        out = dec.predict(z)
    finally:
        sr.set_precision("bf16x3")                         # back to the library's default
    assert not sr.tc_error()
    assert out.shape == ref.shape == (2, 400, 400, 1)
    scale = np.abs(ref).max()
    err = np.abs(out - ref)
    assert err.max() <= 4e-3 * scale, (err.max(), scale, np.unravel_index(err.argmax(), err.shape))
    # borders and tile seams included: rows / columns at 0, 127|128, 399 are no worse than the interior
    for sl in (np.s_[:, 0], np.s_[:, 399], np.s_[:, :, 0], np.s_[:, :, 399], np.s_[:, :, 127:129], np.s_[:, 7:9]):
        assert err[sl].max() <= 4e-3 * scale


@pytest.mark.parametrize("B", [1, 5, 130])
def test_fused_decoder_tail_equals_the_two_launches(B):
    """Split-operand path: the fused tail (last ConvT + swish + final 3x3 conv in one kernel, the 400x400x8 activation in
    shared memory; patches of 14x14 input pixels, ragged at the right/bottom edge, zero padding at every image border)
    against the same two layers as separate launches: same operands, MMA order and FMA order, hence the same bits."""
    from srcfd import sr
    dec = sr.synthetic_decoder(0)
    z = np.random.default_rng(100 + B).standard_normal((B, 50)).astype(np.float32)
    try:
        sr.set_precision("bf16x3_unfused")
        ref = dec.predict(z)
        sr.set_precision("bf16x3")
        out = dec.predict(z)
    finally:
        sr.set_precision("bf16x3")
    assert not sr.tc_error()
    assert np.array_equal(out, ref), (np.abs(out - ref).max(), np.unravel_index(np.abs(out - ref).argmax(), out.shape))
    assert np.abs(ref).max() > 1e-3


def test_super_resolve_device_pipeline(golden_dir):
    """srcfd_sr_super_resolve: statistics blend, standardisation, inverse standardisation and the NaN/Inf guard on the
    device, many fields per call -- against the reference's host statements around the same (GPU) networks."""
    from srcfd import sr
    from srcfd.workflow import standardize_with_stats, inverse_standardize
    enc = sr.load_model(os.path.join(golden_dir, "encoder10_multiBC.h5"))
    dec = sr.synthetic_decoder(0)
    rng = np.random.default_rng(11)
    B = 7
    x = (0.3 * rng.standard_normal((B, 10, 10)) + rng.uniform(-1, 1, (B, 1, 1))).astype(np.float32)
    stats = np.stack([rng.uniform(-0.2, 0.2, B), rng.uniform(0.1, 0.5, B), rng.uniform(-0.1, 0.1, B), rng.uniform(0.2, 0.6, B)], axis=1)
    x[3] = 0.25                                            # constant field: its own std is 0 -> max(std, 1e-8)
    stats[4, 1] = 0.0                                      # zero training std -> 1e-8 when not blended
    stats[5, 3] = np.inf                                   # overflow in the inverse standardisation -> the guard writes zeros
    for adaptive in (False, True):
        got = sr.super_resolve(enc, dec, x, stats, adaptive, 0.3)
        assert got.shape == (B, 400, 400) and got.dtype == np.float32
        for b in range(B):
            m, s = stats[b, 0], stats[b, 1]
            if adaptive:
                m = 0.7 * m + 0.3 * np.mean(x[b]); s = 0.7 * s + 0.3 * max(np.std(x[b]), 1e-8)
            xn = standardize_with_stats(x[b], np.float32(m), np.float32(s if s != 0 else 1e-8))
            y = sr.predict(enc, dec, xn[None, ..., None])[0, ..., 0]
            ref = inverse_standardize(y, np.float32(stats[b, 2]), np.float32(stats[b, 3]))
            ref = np.nan_to_num(ref, nan=0.0, posinf=0.0, neginf=0.0) if not np.isfinite(ref).all() else ref
            if b == 5:
                assert np.isfinite(got[b]).all() and np.count_nonzero(got[b]) < got[b].size // 2
                continue
            if b == 4 and not adaptive:
                assert np.isfinite(got[b]).all()           # 1e-8 divisor: huge but finite inputs, nothing to compare closely
                continue
            np.testing.assert_allclose(got[b], ref, rtol=2e-4, atol=2e-4 * max(1.0, float(np.max(np.abs(ref)))))
    assert sr.launch_count() > 0


@pytest.mark.parametrize("B", [1, 5, 130])
def test_decoder_split_operand_tensor_cores_meet_the_fp32_bar(B):
    """precision 'bf16x3': every ConvT on tcgen05 as three bf16 MMAs per K-step (a_hi*w_hi + a_hi*w_lo + a_lo*w_hi) with
    fp32 activations, epilogues and final conv -- held to the SAME tolerance as the fp32 CUDA-core path
    (rtol 1e-4 / atol 5e-5 against the numpy restatement), and equal to that path to the same bar."""
    from srcfd import sr
    dec = sr.synthetic_decoder(0)
    z = np.random.default_rng(100 + B).standard_normal((B, 50)).astype(np.float32)
    try:
        sr.set_precision("bf16x3")
        out = dec.predict(z)
        sr.set_precision("fp32")
        ref32 = dec.predict(z)
    finally:
        sr.set_precision("bf16x3")                         # the library's default
    assert not sr.tc_error()
    nref = min(B, 2)
    np.testing.assert_allclose(out[:nref], S.decoder_forward(z[:nref], dec.weights), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(out, ref32, rtol=1e-4, atol=5e-5)
    if B > 128:                                            # the second 128-sample chunk, and a partial last tile per layer
        np.testing.assert_allclose(out[129], ref32[129], rtol=1e-4, atol=5e-5)
