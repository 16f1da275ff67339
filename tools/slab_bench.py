"""BASELINE configs[3] microbenchmark: Jacobi-order pressure relaxation on a large grid split into row slabs over N GPUs
(one process per GPU, NCCL halo exchange once per H-sweep pass, one all-reduce of the per-sweep residual sums per pass).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_bench.py [n] [sweeps]
Prints one JSON line on rank 0: strong scaling (the grid is fixed, ranks split it)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sr-for-cfd_b200"))
import numpy as np
import torch
import torch.distributed as dist
from srcfd.slab import GpuSlab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rng = np.random.default_rng(0)
Var = np.zeros((3, n + 2, n + 2)); Var[2] = rng.uniform(-1, 1, (n + 2, n + 2))
Ff = 1e-3 * rng.uniform(-1, 1, (4, n + 2, n + 2))
slab = GpuSlab(n, n, 1.0 / n, 1.0 / n, 1e-3, 1.0, Var, Ff, world, rank, device=local,
               passes_per_exchange=int(os.environ.get("SLAB_M", "8")))
del Var, Ff
slab.solve(tol=0.0, max_iter=2 * max(1, slab.nsw_max))          # warm-up (also builds the right-hand side)
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
done, rms = slab.solve(tol=0.0, max_iter=sweeps)
slab.h.synchronize(); torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
if rank == 0:
    print(json.dumps({"metric": "Jacobi pressure relaxation, slab decomposition", "grid": [n, n], "n_gpus": world, "sweeps": done,
                      "sweeps_per_pass": slab.nsw_max, "halo_rows": slab.part.halo, "seconds": dt, "us_per_sweep": 1e6 * dt / done,
                      "value": n * n * done / dt / 1e9, "unit": "GLUP/s", "scaling": "strong",
                      "algorithmic_GBs": 24 * n * n * done / dt / 1e9, "last_rms": rms,
                      "halo_bytes_per_pass_per_neighbour": slab.part.halo * (n + 2) * 8}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
