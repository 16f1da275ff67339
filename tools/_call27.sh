cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/c27_bench8.json 2> gpurun_out/c27_bench8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c27_bench8.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
sl=d['slab']; print('pressure', sl['pressure']['value'], sl['pressure'].get('strong_scaling_efficiency_same_run'), sl['pressure']['slab_parity'])
print('outer', sl['outer_iterations']['ms_per_iteration'], sl['outer_iterations']['slab_parity'], sl.get('slab_parity'))
print('developed', sl.get('outer_iterations_developed'))
print('weak', sl.get('pressure_weak',{}).get('value'), 'ensemble', d['ensemble']['value'])
PY
tail -3 gpurun_out/c27_bench8.err
