"""BASELINE.json configs[4]: SR conv-AE decoder throughput, batch 1024 latents -> 1024 x (400,400,1) fields, 1 x B200.
Inputs and outputs resident in HBM (torch tensors used only as device memory); CUDA-event time inside the library."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
import torch
from srcfd import sr

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1024); ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
B = args.batch
FLOP_PER_SAMPLE = 2 * (50 * 36864 + 144 * 256 * 9 * 128 + 625 * 128 * 256 + 2500 * 64 * 128 + 10000 * 32 * 64 + 40000 * 16 * 32 + 160000 * 72)
dec = sr.synthetic_decoder(0)
z = torch.randn(B, 50, device="cuda", dtype=torch.float32, generator=torch.Generator("cuda").manual_seed(0))
out = torch.empty(B, 400, 400, 1, device="cuda", dtype=torch.float32)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
res = {}
for mode in ("fp32", "bf16x3_unfused", "bf16x3", "bf16"):
    sr.set_precision(mode)
    for _ in range(3):
        sr.decode_device(dec, z.data_ptr(), min(B, 64), out.data_ptr())
    ms = min(sr.decode_device(dec, z.data_ptr(), B, out.data_ptr()) for _ in range(args.reps))
    torch.cuda.synchronize()
    res[mode] = {"ms": ms, "samples_per_s": B / (ms * 1e-3), "tflops": B * FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12,
                 "output_write_gbs": B * 640000 / (ms * 1e-3) / 1e9}
sr.set_precision("bf16x3")
print(json.dumps({"metric": "SR decoder inference throughput", "unit": "samples/s", "batch": B, "value": res["bf16"]["samples_per_s"],
                  "dtype": "bf16 operands / f32 accumulate (tcgen05) for the 5 ConvT layers and the final 3x3 conv; f32 CUDA cores for Dense",
                  "flop_per_sample": FLOP_PER_SAMPLE, "paths": res, "tc_error": sr.tc_error(),
                  "roofline": {"bound": "tensor", "achieved": res["bf16"]["tflops"], "peak": peaks.get("bf16_tflops_sustained"), "unit": "TFLOP/s",
                               "frac": res["bf16"]["tflops"] / peaks.get("bf16_tflops_sustained", 1400.0),
                               "hbm_output_frac": res["bf16"]["output_write_gbs"] / peaks["hbm_gbs"]},
                  "data": "synthetic (Glorot seed-0 decoder weights, N(0,1) latents seed 0)"}))
