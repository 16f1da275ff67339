"""4096^2 JACOBI momentum sweeps, upwind and QUICK, on face fluxes produced by the library's own kernels: two sweeps per
pass (k_slab_sweep2, the default) against one sweep per launch (k_slab_sweep, SRCFD_SLAB_SWEEP2=0), and a scan of the row
chunks per strip (SRCFD_SWEEP2_CHUNKS); GLUP/s and the fraction of the measured HBM copy bandwidth at SURVEY 8d's 40
algorithmic bytes per cell update.  The knobs are read when a slab is configured, so one process can compare them.
SRCFD_SLAB_FOUR_FACES=1 reads all four flux planes; SRCFD_LIB selects another build of the library."""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import bench
from srcfd import slab, _capi as capi
n, out = int(os.environ.get("PROBE_N", "4096")), {}
peak, _ = bench.peaks()
variants = [("two_per_pass", {"SRCFD_SLAB_SWEEP2": "1"}), ("one_per_launch", {"SRCFD_SLAB_SWEEP2": "0"})]
for ch in os.environ.get("PROBE_CHUNKS", "8,12,24,34,64").split(","):
    if ch:
        variants.append((f"two_per_pass_chunks{ch}", {"SRCFD_SLAB_SWEEP2": "1", "SRCFD_SWEEP2_CHUNKS": ch}))
if os.environ.get("PROBE_SFN"):
    variants.append(("one_per_launch_south_from_north", {"SRCFD_SLAB_SWEEP2": "0", "SRCFD_SLAB_SOUTH_FROM_NORTH": "1"}))
for pf in os.environ.get("PROBE_PF", "").split(","):
    if pf:
        variants.append((f"two_per_pass_pf{pf}", {"SRCFD_SLAB_SWEEP2": "1", "SRCFD_SWEEP2_PF": pf}))
for vname, env in variants:
  for key in ("SRCFD_SLAB_SWEEP2", "SRCFD_SWEEP2_CHUNKS", "SRCFD_SWEEP2_PF", "SRCFD_SLAB_SOUTH_FROM_NORTH"):
      os.environ.pop(key, None)
  os.environ.update(env)
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
      V = np.zeros((3, n + 2, n + 2))
      s.h.download(Var=V)
      out[f"{vname}/{name}"] = {"sweeps": sw, "ms": round(ms, 3), "glups": round(n * n * sw / ms / 1e6, 1), "frac40": round(40 * n * n * sw / ms / 1e6 / peak, 3),
                                "u_sha1": hashlib.sha1(V[0].tobytes()).hexdigest()[:12]}
      del V
      s.close()
for name in ("upwind", "quick"):
    ref = out[f"one_per_launch/{name}"]["u_sha1"]
    out[f"fields_equal/{name}"] = all(v["u_sha1"] == ref for k, v in out.items() if k.endswith("/" + name) and isinstance(v, dict))
print(json.dumps(out))
