"""One 4096^2 upwind momentum solve (32 JACOBI sweeps, library-produced fluxes) for an ncu capture of k_slab_sweep2 /
k_slab_sweep (SRCFD_SLAB_SWEEP2=0); PROBE_QUICK=1 for the QUICK stencil."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import bench
from srcfd import slab, _capi as capi
n = int(os.environ.get("PROBE_N", "4096"))
quick = os.environ.get("PROBE_QUICK", "0") == "1"
s = slab.GpuSlab(bench._ldc_params(n, 0, 32, 0.0, scheme_quick=quick), 1, 0)
Var, Ff = bench._synthetic_rows(n, 0, n + 1)
s.h.upload(Var=Var, VarOld=Var)
for k in range(3):
    s.h.k_apply_bc(k)
s.h.k_linear_interpolation(); s.h.k_update_flux()
sw, rms = slab.solve_momentum([s], 0, capi.SCHEME_QUICK if quick else capi.SCHEME_UPWIND)
s.h.synchronize()
print("sweeps", sw, "rms", rms)
s.close()
