"""BASELINE configs[3] microbenchmark: JACOBI-order pressure relaxation (and whole outer iterations) of a large
lid-driven-cavity grid split into row slabs over N GPUs -- one process per GPU, halo rows and residual sums pushed by
the kernels through cudaIpc-mapped peer memory (csrc/slab.cuh), block loop inside libsrcfd.
  python tools/slab_bench.py [n] [sweeps] [outer]                                                      (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_bench.py [n] [sweeps] [outer]
Prints one JSON line on rank 0: strong scaling (the grid is fixed, the ranks split it)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
import torch
import torch.distributed as dist
from srcfd import slab, _capi as capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 0
halo = int(os.environ.get("SLAB_HALO", "16"))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))


def params(inner_max, scheme=capi.SCHEME_QUICK, tol=0.0):
    p = capi.Params()
    p.nx = p.ny = n
    p.dx = p.dy = 1.0 / n
    p.volp = p.dx * p.dy
    p.dt, p.nu, p.rho = 1e-3, 1.0 / 1000.0, 1.0
    p.scheme = scheme
    for k in range(3):
        for s in range(4):
            p.bc_types[k][s] = 1 if k == 2 else 0
    p.bc_values[0][2] = 1.0                                   # lid
    p.inner_tol, p.inner_max, p.sweep_order, p.device = tol, inner_max, capi.ORDER_JACOBI, local
    return p


def maxtime(ms):
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


s = slab.GpuSlab(params(sweeps), world, rank, halo=halo)
slab.attach_distributed(s)
g0, g1 = s.part.global_rows()
rng = np.random.default_rng(0)
Var = np.zeros((3, g1 - g0 + 1, n + 2)); Ff = np.zeros((4, g1 - g0 + 1, n + 2))
full = np.random.default_rng(0)
# rows are generated per global row so that every world size sees the same field
for r in range(g0, g1 + 1):
    rr = np.random.default_rng(1000 + r)
    Var[2, r - g0] = rr.uniform(-1, 1, n + 2)
    Ff[:, r - g0] = 1e-3 * rr.uniform(-1, 1, (4, n + 2))
s.h.upload(Var=Var, Ff=Ff)
del Var, Ff
out = {"metric": "JACOBI pressure relaxation, slab decomposition (peer-memory halos, in-library block loop)", "grid": [n, n],
       "n_gpus": world, "halo_rows": s.part.halo, "scaling": "strong"}
slab.solve_pressure([s])                                      # warm-up
torch.cuda.synchronize()
if world > 1: dist.barrier()
i0 = s.info()
s.h.timer_start()
done, rms = slab.solve_pressure([s])
ms = maxtime(s.h.timer_stop())
i1 = s.info()
out.update({"sweeps": done, "ms": ms, "us_per_sweep": 1e3 * ms / done, "value": n * n * done / ms / 1e6, "unit": "GLUP/s",
            "algorithmic_GBs": 24 * n * n * done / ms / 1e6, "last_rms": rms,
            "exchanges": i1["exchanges"] - i0["exchanges"], "halo_bytes_pushed_rank0": i1["halo_bytes"] - i0["halo_bytes"]})
if outer > 0:
    s2 = slab.GpuSlab(params(1000, tol=1e-6), world, rank, halo=halo)
    slab.attach_distributed(s2)
    s2.h.initialize_fields(True)
    slab.step([s2], 1, (0.0, 0.0, 0.0))
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    st0 = s2.h.status()
    s2.h.timer_start()
    slab.step([s2], outer, (0.0, 0.0, 0.0))
    ms2 = maxtime(s2.h.timer_stop())
    st1 = s2.h.status()
    sw = (st1["total_sweeps"] - st0["total_sweeps"]).astype(np.int64)
    out["outer"] = {"iterations": outer, "ms_per_iteration": ms2 / outer, "inner_sweeps": sw.tolist(),
                    "value": n * n * float(sw.sum()) / ms2 / 1e6, "unit": "GLUP/s", "rms": st1["rms"].tolist(),
                    "kernel_stats": s2.kernel_stats()}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
