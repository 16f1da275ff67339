"""One bf16 decode of a batch (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import torch
from srcfd import sr
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dec = sr.synthetic_decoder(0)
z = torch.randn(B, 50, device="cuda", dtype=torch.float32, generator=torch.Generator("cuda").manual_seed(0))
out = torch.empty(B, 400, 400, 1, device="cuda", dtype=torch.float32)
sr.set_precision("bf16")
for _ in range(2):
    ms = sr.decode_device(dec, z.data_ptr(), B, out.data_ptr())
print(B, ms, "ms ->", B / ms * 1e3, "samples/s")
