"""4096^2 JACOBI momentum sweeps (k_slab_sweep), upwind and QUICK, on face fluxes produced by the library's own kernels;
GLUP/s and the fraction of the measured HBM copy bandwidth at SURVEY 8d's 40 algorithmic bytes per cell update.
SRCFD_SLAB_FOUR_FACES=1 reads all four flux planes; SRCFD_LIB selects another build of the library."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import bench
from srcfd import slab, _capi as capi
n, out = 4096, {}
peak, _ = bench.peaks()
for scheme, name in ((False, "upwind"), (True, "quick")):
    s = slab.GpuSlab(bench._ldc_params(n, 0, 32, 0.0, scheme_quick=scheme), 1, 0)
    Var, Ff = bench._synthetic_rows(n, 0, n + 1)
    s.h.upload(Var=Var, VarOld=Var)
    for k in range(3):
        s.h.k_apply_bc(k)
    s.h.k_linear_interpolation(); s.h.k_update_flux()
    sc = capi.SCHEME_QUICK if scheme else capi.SCHEME_UPWIND
    slab.solve_momentum([s], 0, sc)
    s.h.synchronize()
    s.h.timer_start()
    sw, _ = slab.solve_momentum([s], 0, sc)
    ms = s.h.timer_stop()
    out[name] = {"sweeps": sw, "ms": round(ms, 3), "glups": round(n * n * sw / ms / 1e6, 1), "frac40": round(40 * n * n * sw / ms / 1e6 / peak, 3)}
    s.close()
print(json.dumps(out))
