"""The schedule of k_slab_sweep2 (two JACOBI momentum sweeps per pass, csrc/slab.cuh) modelled on the CPU: strip / chunk
tiling, the register windows and the input ring indexed as the kernel's unrolled step loop indexes them, QUICK's
out-of-plane reads at both levels, paired fluxes -- against two JACOBI sweeps of the oracle.  (tools/sim_sweep2.py holds
the model; the kernel itself is checked on the GPU in tests/test_gpu_slab.py.)"""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("quick", [False, True])
@pytest.mark.parametrize("paired", [False, True])
def test_two_sweeps_per_pass_schedule_equals_two_oracle_sweeps(quick, paired):
    import sim_sweep2
    for nx, ny, RB in ((96, 50, 32), (33, 28, 7), (5, 3, 5), (64, 90, 13), (41, 57, 41)):
        for kpl in (0, 1):
            err, same = sim_sweep2.check(nx, ny, quick, RB, kpl=kpl, paired=paired, seed=nx + kpl)
            assert same, (quick, paired, nx, ny, RB, kpl, err)
