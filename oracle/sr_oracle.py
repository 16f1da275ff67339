"""CPU restatement of the SR autoencoder inference path (encoder_10 + decoder_400).

TEST INFRASTRUCTURE ONLY.  Parity status: **UNPINNED for the decoder** -- the reference tree ships
no decoder weights (.MISSING_LARGE_BLOBS:29-34), no stored encoder/decoder input-output pair, and
TensorFlow/Keras (requirements.txt:6-7, unpinned; the .h5 metadata says Keras 3.8.0, backend
tensorflow) is not installed, so nothing the reference holds can check this file.  It restates the
published Keras layer semantics for the architecture in sr-ae-conv.ipynb (cell lines 162-169 and
277-287) and the call sites PyCFD_ML_accelerated.py:839-876, and is cross-checked against an
independent implementation built from torch's CPU conv ops (tests/test_sr_oracle.py).  The encoder
weights are real (committed .h5); decoder weights are synthetic.

Keras semantics restated here
  Conv2D 'same', stride s   out = ceil(in/s); total pad = max((out-1)*s + k - in, 0), the odd unit goes to
                            the END (bottom/right); kernel layout (kh, kw, Cin, Cout); cross-correlation.
  Conv2DTranspose 'valid'   out = (in-1)*s + k; out[y*s+ky, x*s+kx, co] += in[y, x, ci] * W[ky, kx, co, ci];
                            kernel layout (kh, kw, Cout, Cin).
  Dense                     x @ W + b with W (in, out).   Flatten / Reshape: row-major (h, w, c).
  swish (saved as "silu")   x * sigmoid(x).               Everything float32, NHWC.
"""
from __future__ import annotations

import numpy as np


def swish(x):
    x = x.astype(np.float32)
    return (x / (np.float32(1.0) + np.exp(-x))).astype(np.float32)


def dense(x, W, b):
    return (x.astype(np.float32) @ W.astype(np.float32) + b.astype(np.float32)).astype(np.float32)


def conv2d_same(x, W, b, stride=1):
    """x (B,H,Wd,Cin), W (kh,kw,Cin,Cout)."""
    B, H, Wd, Cin = x.shape
    kh, kw, _, Cout = W.shape
    oh, ow = -(-H // stride), -(-Wd // stride)
    ph = max((oh - 1) * stride + kh - H, 0)
    pw = max((ow - 1) * stride + kw - Wd, 0)
    xp = np.zeros((B, H + ph, Wd + pw, Cin), dtype=np.float32)
    xp[:, ph // 2: ph // 2 + H, pw // 2: pw // 2 + Wd] = x
    out = np.zeros((B, oh, ow, Cout), dtype=np.float32)
    for ky in range(kh):
        for kx in range(kw):
            patch = xp[:, ky: ky + (oh - 1) * stride + 1: stride, kx: kx + (ow - 1) * stride + 1: stride, :]
            out += np.tensordot(patch, W[ky, kx].astype(np.float32), axes=([3], [0]))
    return (out + b.astype(np.float32)).astype(np.float32)


def conv2d_transpose_valid(x, W, b, stride):
    """x (B,H,Wd,Cin), W (kh,kw,Cout,Cin)."""
    B, H, Wd, Cin = x.shape
    kh, kw, Cout, _ = W.shape
    oh, ow = (H - 1) * stride + kh, (Wd - 1) * stride + kw
    out = np.zeros((B, oh, ow, Cout), dtype=np.float32)
    for ky in range(kh):
        for kx in range(kw):
            contrib = np.tensordot(x, W[ky, kx].astype(np.float32), axes=([3], [1]))   # (B,H,Wd,Cout)
            out[:, ky: ky + (H - 1) * stride + 1: stride, kx: kx + (Wd - 1) * stride + 1: stride, :] += contrib
    return (out + b.astype(np.float32)).astype(np.float32)


def encoder_forward(x, w):
    """encoder_10 (sr-ae-conv.ipynb cell 162-169).  x: (B,10,10,1) float32 -> (B,50)."""
    h = swish(conv2d_same(x.astype(np.float32), w["conv2d/kernel"], w["conv2d/bias"], stride=2))     # (B,5,5,64)
    h = swish(conv2d_same(h, w["conv2d_1/kernel"], w["conv2d_1/bias"], stride=1))                      # (B,5,5,128)
    h = h.reshape(h.shape[0], -1)                                                                      # (B,3200)
    h = swish(dense(h, w["dense/kernel"], w["dense/bias"]))                                            # (B,128)
    return dense(h, w["latent_vector/kernel"], w["latent_vector/bias"])                                # (B,50)


DECODER_LAYERS = [  # name, kind, kernel shape (Keras layout)
    ("dense", "dense", (50, 12 * 12 * 256)),
    ("conv2d_transpose", "convT", (3, 3, 128, 256)),
    ("conv2d_transpose_1", "convT", (2, 2, 64, 128)),
    ("conv2d_transpose_2", "convT", (2, 2, 32, 64)),
    ("conv2d_transpose_3", "convT", (2, 2, 16, 32)),
    ("conv2d_transpose_4", "convT", (2, 2, 8, 16)),
    ("output_image_400", "conv", (3, 3, 8, 1)),
]


def decoder_forward(z, w, return_all=False):
    """decoder_400 (sr-ae-conv.ipynb cell 277-287).  z: (B,50) -> (B,400,400,1)."""
    acts = []
    h = swish(dense(z.astype(np.float32), w["dense/kernel"], w["dense/bias"])).reshape(-1, 12, 12, 256)
    acts.append(h)
    for i, name in enumerate(["conv2d_transpose", "conv2d_transpose_1", "conv2d_transpose_2", "conv2d_transpose_3",
                              "conv2d_transpose_4"]):
        h = swish(conv2d_transpose_valid(h, w[f"{name}/kernel"], w[f"{name}/bias"], stride=2))
        acts.append(h)
    out = conv2d_same(h, w["output_image_400/kernel"], w["output_image_400/bias"], stride=1)
    return (out, acts) if return_all else out


# ---- independent implementation from torch CPU ops (cross-check only) ------------------------------
def torch_encoder_forward(x, w):
    import torch
    import torch.nn.functional as F
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    h = t(x).permute(0, 3, 1, 2)                                    # NCHW
    h = F.pad(h, (0, 1, 0, 1))                                      # TF 'same', stride 2 on 10: pad end only
    h = F.silu(F.conv2d(h, t(w["conv2d/kernel"]).permute(3, 2, 0, 1), t(w["conv2d/bias"]), stride=2))
    h = F.silu(F.conv2d(h, t(w["conv2d_1/kernel"]).permute(3, 2, 0, 1), t(w["conv2d_1/bias"]), padding=1))
    h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)               # flatten in (h, w, c) order
    h = F.silu(h @ t(w["dense/kernel"]) + t(w["dense/bias"]))
    return (h @ t(w["latent_vector/kernel"]) + t(w["latent_vector/bias"])).numpy()


def torch_decoder_forward(z, w):
    import torch
    import torch.nn.functional as F
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    h = F.silu(t(z) @ t(w["dense/kernel"]) + t(w["dense/bias"])).reshape(-1, 12, 12, 256).permute(0, 3, 1, 2)
    for name in ["conv2d_transpose", "conv2d_transpose_1", "conv2d_transpose_2", "conv2d_transpose_3", "conv2d_transpose_4"]:
        W = t(w[f"{name}/kernel"]).permute(3, 2, 0, 1)               # (Cin, Cout, kh, kw)
        h = F.silu(F.conv_transpose2d(h, W, t(w[f"{name}/bias"]), stride=2))
    W = t(w["output_image_400/kernel"]).permute(3, 2, 0, 1)
    h = F.conv2d(h, W, t(w["output_image_400/bias"]), padding=1)
    return h.permute(0, 2, 3, 1).numpy()
