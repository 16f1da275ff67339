"""SR autoencoder inference on the GPU (encoder_10 + decoder_400) behind Keras-like model objects.

    encoder = load_model("vanilla_encoder10_to_400_*.h5")     # real committed weights, read with h5lite
    decoder = load_model("vanilla_decoder400_from_10_*.h5")   # same layout (the reference tree lacks these files)
    decoder = synthetic_decoder(seed=0)                        # Glorot-uniform stand-in used by benchmarks/tests
    y = SuperResolutionAE(encoder, decoder).predict(x)         # (B,10,10,1) -> (B,400,400,1), float32

replaces `tf.keras.models.load_model(...)` / `.predict` at PyCFD_ML_accelerated.py:831-858.  The layers run
in libsrcfd (sr-for-cfd_b200/csrc/sr.cu) through the C ABI; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi as capi
from . import h5lite

ENCODER_LAYERS = ["conv2d", "conv2d_1", "dense", "latent_vector"]
DECODER_LAYERS = ["dense", "conv2d_transpose", "conv2d_transpose_1", "conv2d_transpose_2", "conv2d_transpose_3",
                  "conv2d_transpose_4", "output_image_400"]
DECODER_SHAPES = {   # Keras kernel shapes (sr-ae-conv.ipynb cell 277-287)
    "dense": (50, 12 * 12 * 256), "conv2d_transpose": (3, 3, 128, 256), "conv2d_transpose_1": (2, 2, 64, 128),
    "conv2d_transpose_2": (2, 2, 32, 64), "conv2d_transpose_3": (2, 2, 16, 32), "conv2d_transpose_4": (2, 2, 8, 16),
    "output_image_400": (3, 3, 8, 1)}
ENCODER_SHAPES = {"conv2d": (3, 3, 1, 64), "conv2d_1": (3, 3, 64, 128), "dense": (3200, 128), "latent_vector": (128, 50)}

_fp = C.POINTER(C.c_float)
_ctx = {}
default_device = 0


def _sr_check(rc):
    if rc != 0:
        raise capi.SrcfdError(f"libsrcfd SR error {rc}: {capi.lib().srcfd_sr_last_error().decode('utf8', 'replace')}")


def _context(device=None):
    device = default_device if device is None else device
    if device not in _ctx:
        L = capi.lib()
        L.srcfd_sr_last_error.restype = C.c_char_p
        h = C.c_void_p()
        _sr_check(L.srcfd_sr_create(C.c_int(device), C.byref(h)))
        _ctx[device] = {"h": h, "enc": None, "dec": None}
    return _ctx[device]


def read_keras_weights(path: str) -> dict:
    """{'<layer>/kernel': array, '<layer>/bias': array} from a Keras legacy-H5 file (model_weights/<l>/<l>/...)."""
    root = h5lite.read_h5(path)
    mw = root["model_weights"] if "model_weights" in root else root
    out = {}
    for lname, grp in mw.items():
        if not isinstance(grp, h5lite.Group):
            continue
        for path_, node in grp.visit():
            if isinstance(node, h5lite.Dataset):
                leaf = path_.strip("/").split("/")[-1].split(":")[0]
                out[f"{lname}/{leaf}"] = np.ascontiguousarray(node.data, dtype=np.float32)
    return out


def glorot_decoder_weights(seed: int = 0) -> dict:
    """Glorot-uniform kernels (Keras fan rules) and small uniform biases for decoder_400."""
    rng = np.random.default_rng(seed)
    w = {}
    for name in DECODER_LAYERS:
        shape = DECODER_SHAPES[name]
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = rf * shape[-2], rf * shape[-1]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        w[f"{name}/kernel"] = rng.uniform(-lim, lim, shape).astype(np.float32)
        nb = shape[-1] if "transpose" not in name else shape[-2]
        w[f"{name}/bias"] = rng.uniform(-0.05, 0.05, nb).astype(np.float32)
    return w


def resolve_layers(weights: dict, layers, shapes) -> dict:
    """Map the layers of a weight file onto the canonical layer list BY KERNEL SHAPE, not by name.

    Keras numbers auto-named layers per session: a decoder built after the encoder in the same notebook
    (sr-ae-conv.ipynb) gets 'dense_1', 'conv2d_transpose_5', ...; the kernel shapes of both networks are pairwise
    distinct, so each canonical layer is the one unused layer of the file whose kernel has its shape (an exact name
    match wins when several qualify).  Returns {canonical/kernel|bias: array}; ValueError names what is missing."""
    have = [k[:-len("/kernel")] for k in weights if k.endswith("/kernel")]
    used, out = set(), {}
    for name in layers:
        want = tuple(shapes[name])
        cands = [l for l in have if l not in used and tuple(weights[f"{l}/kernel"].shape) == want]
        if name in cands:
            pick = name
        elif len(cands) >= 1:
            pick = cands[0]
        else:
            got = {l: tuple(weights[f"{l}/kernel"].shape) for l in have}
            if name in got:
                raise ValueError(f"{name}/kernel has shape {got[name]}, expected {want}")
            raise ValueError(f"no layer with a {want} kernel for '{name}' in the weight file (layers: {got})")
        if f"{pick}/bias" not in weights:
            raise ValueError(f"layer '{pick}' has a kernel but no bias")
        used.add(pick)
        out[f"{name}/kernel"], out[f"{name}/bias"] = weights[f"{pick}/kernel"], weights[f"{pick}/bias"]
    return out


class _Model:
    kind = ""

    def __init__(self, weights: dict, device=None):
        layers, shapes = (ENCODER_LAYERS, ENCODER_SHAPES) if self.kind == "enc" else (DECODER_LAYERS, DECODER_SHAPES)
        weights = resolve_layers(weights, layers, shapes)
        self.weights = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
        self.device = default_device if device is None else device
        self._layers = layers

    def _bind(self):
        ctx = _context(self.device)
        if ctx[self.kind] is not self:
            n = len(self._layers)
            ks = (_fp * n)(*[self.weights[f"{l}/kernel"].ctypes.data_as(_fp) for l in self._layers])
            bs = (_fp * n)(*[self.weights[f"{l}/bias"].ctypes.data_as(_fp) for l in self._layers])
            fn = capi.lib().srcfd_sr_set_encoder if self.kind == "enc" else capi.lib().srcfd_sr_set_decoder
            _sr_check(fn(ctx["h"], ks, bs))
            ctx[self.kind] = self
        return ctx["h"]

    def predict(self, x, verbose=0):
        return self(x)


class Encoder(_Model):
    """encoder_10: (B,10,10,1) -> (B,50)."""
    kind = "enc"

    def __call__(self, x, training=False):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 10, 10, 1)
        z = np.empty((x.shape[0], 50), dtype=np.float32)
        _sr_check(capi.lib().srcfd_sr_encode(self._bind(), x.ctypes.data_as(_fp), C.c_int(x.shape[0]), z.ctypes.data_as(_fp)))
        return z


class Decoder(_Model):
    """decoder_400: (B,50) -> (B,400,400,1)."""
    kind = "dec"

    def __call__(self, z, training=False):
        z = np.ascontiguousarray(z, dtype=np.float32).reshape(-1, 50)
        out = np.empty((z.shape[0], 400, 400, 1), dtype=np.float32)
        _sr_check(capi.lib().srcfd_sr_decode(self._bind(), z.ctypes.data_as(_fp), C.c_int(z.shape[0]), out.ctypes.data_as(_fp)))
        return out


def load_model(path_or_model, compile=False):
    """tf.keras.models.load_model stand-in for the two SR networks (or pass an Encoder/Decoder through)."""
    if isinstance(path_or_model, _Model):
        return path_or_model
    if not os.path.exists(path_or_model):
        raise OSError(f"No file or directory found at {path_or_model}")
    w = read_keras_weights(path_or_model)
    kshapes = {tuple(v.shape) for k, v in w.items() if k.endswith("/kernel")}
    if ENCODER_SHAPES["latent_vector"] in kshapes and ENCODER_SHAPES["conv2d"] in kshapes:
        return Encoder(w)
    if DECODER_SHAPES["output_image_400"] in kshapes and DECODER_SHAPES["dense"] in kshapes:
        return Decoder(w)
    raise ValueError(f"{path_or_model}: not an encoder_10 / decoder_400 weight file")


def synthetic_decoder(seed: int = 0) -> Decoder:
    return Decoder(glorot_decoder_weights(seed))


def predict(encoder: Encoder, decoder: Decoder, x) -> np.ndarray:
    """SuperResolutionAE.call in one library call (latents never leave the device)."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 10, 10, 1)
    assert encoder.device == decoder.device
    encoder._bind(); h = decoder._bind()
    out = np.empty((x.shape[0], 400, 400, 1), dtype=np.float32)
    _sr_check(capi.lib().srcfd_sr_predict(h, x.ctypes.data_as(_fp), C.c_int(x.shape[0]), out.ctypes.data_as(_fp)))
    return out


def super_resolve(encoder: Encoder, decoder: Decoder, x, stats, adaptive: bool = False, blend: float = 0.3) -> np.ndarray:
    """The per-field pipeline of ml_super_resolution for B fields in one library call (device-side statistics blend,
    standardisation, encoder, decoder, inverse standardisation, NaN/Inf guard).  x (B,10,10); stats (B,4) =
    {mean_lr, std_lr, mean_hr, std_hr}; returns (B,400,400) float32."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 10, 10)
    st = np.ascontiguousarray(stats, dtype=np.float64).reshape(-1, 4)
    assert st.shape[0] == x.shape[0] and encoder.device == decoder.device
    encoder._bind(); h = decoder._bind()
    out = np.empty((x.shape[0], 400, 400), dtype=np.float32)
    _sr_check(capi.lib().srcfd_sr_super_resolve(h, x.ctypes.data_as(_fp), C.c_int(x.shape[0]), st.ctypes.data_as(C.POINTER(C.c_double)),
                                                C.c_int(int(bool(adaptive))), C.c_double(float(blend)), out.ctypes.data_as(_fp)))
    return out


def decode_device(decoder: Decoder, z_dev_ptr: int, B: int, out_dev_ptr: int) -> float:
    """Decoder on device-resident buffers; returns the CUDA-event time in ms (throughput benchmark)."""
    ms = C.c_double(0.0)
    _sr_check(capi.lib().srcfd_sr_decode_device(decoder._bind(), C.c_uint64(z_dev_ptr), C.c_int(B), C.c_uint64(out_dev_ptr), C.byref(ms)))
    return ms.value


def launch_count(device=None) -> int:
    n = C.c_int64(0)
    _sr_check(capi.lib().srcfd_sr_launch_count(_context(device)["h"], C.byref(n)))
    return n.value


def set_precision(mode: str = "bf16x3", device=None):
    """'fp32' (CUDA cores), 'bf16x3' (tcgen05, split operands: fp32-grade accuracy), 'bf16' (tcgen05, bf16 operands and
    activations: fastest, ~3e-2 of the output range), 'bf16_cc_final' (as bf16 with the CUDA-core final conv; tests),
    'bf16x3_unfused' (as bf16x3 with the last ConvT and the final conv as two launches instead of the fused tail; tests)."""
    _sr_check(capi.lib().srcfd_sr_set_precision(_context(device)["h"], C.c_int({"fp32": 0, "bf16": 1, "bf16_cc_final": 2, "bf16x3": 3, "bf16x3_unfused": 4}[mode])))


def tc_error(device=None) -> bool:
    f = C.c_int(0)
    _sr_check(capi.lib().srcfd_sr_tc_error(_context(device)["h"], C.byref(f)))
    return bool(f.value)


def debug_convT_tc(decoder: Decoder, layer: int, x) -> np.ndarray:
    """One tensor-core ConvT layer (1..4) in isolation: x (B,H,H,Cin) fp32 -> (B,2H,2H,Cout) fp32 (bf16 operands)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, H = x.shape[0], x.shape[1]
    cout = DECODER_SHAPES[DECODER_LAYERS[layer + 1]][2]
    out = np.empty((B, 2 * H, 2 * H, cout), dtype=np.float32)
    _sr_check(capi.lib().srcfd_sr_debug_convT_tc(decoder._bind(), C.c_int(layer), x.ctypes.data_as(_fp), C.c_int(B), out.ctypes.data_as(_fp)))
    return out
