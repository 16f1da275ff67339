"""Host-side layout logic of the decoder's tensor-core kernels, checked without a GPU: the weight image that
k_convT3x3_l1_tc3 streams with cp.async.bulk (every stage the exact shared-memory image its UMMA descriptors read), and the
shared-memory budgets of that kernel and of the fused tail (whose A operand is aliased under the activation tile)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_l1_weight_image_is_the_canonical_umma_layout(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "l1_layout_check")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tests", "csrc", "l1_layout_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok")
