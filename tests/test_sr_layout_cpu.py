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


def test_fused_tail_activation_tile_swizzle_bank_behaviour():
    """tf_chunk (sr_tc.cuh) mirrored: 16-byte chunk of channel half h of pixel (row, col) in the fused tail's 32 x 32 x 8 fp32
    activation tile.  A 16-byte shared-memory access is served a quarter-warp (8 lanes) at a time, one bank group = chunk % 8.
    The epilogue's lanes write pixels two columns (64 B) apart: 4-way conflicts in the linear layout, none with the XOR; the
    conv's lanes read consecutive pixels at every alignment: at most 2-way either way (measured with ncu: 47 M -> 19 M conflict
    wavefronts per pass).  Also: the swizzle is a bijection on the tile."""
    AT = 32

    def tf_chunk(row, col, h):
        pp, q = col >> 1, ((col & 1) << 1) | h
        return (row * AT + 2 * pp) * 2 + (q ^ ((pp >> 1) & 3))

    def linear(row, col, h):
        return (row * AT + col) * 2 + h

    def worst(chunks):
        w = 0
        for s in range(0, len(chunks), 8):
            g = [c % 8 for c in chunks[s:s + 8]]
            w = max(w, max(g.count(x) for x in set(g)))
        return w

    def writer(f):
        return max(worst([f(2 * ((wp * 32 + l) >> 4) + dy, 2 * ((wp * 32 + l) & 15) + dx, h) for l in range(32)])
                   for wp in range(4) for dy in (0, 1) for dx in (0, 1) for h in (0, 1))

    def reader(f):
        return max(worst([f(7 * tq + 1 + r, tx + 1 + kx, h) for tx in range(32)])
                   for tq in range(4) for r in range(9) for kx in range(3) for h in (0, 1))

    assert writer(linear) == 4 and writer(tf_chunk) == 1
    assert reader(tf_chunk) <= 2 and reader(linear) <= 2
    cells = {tf_chunk(r, c, h) for r in range(AT) for c in range(AT) for h in (0, 1)}
    assert cells == set(range(AT * AT * 2))
