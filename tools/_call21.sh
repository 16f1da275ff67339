cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/c21_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c21_tests.log
tail -6 gpurun_out/c21_tests.log
( time timeout 900 python bench.py ) > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c21_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['step_breakdown_ms'])
t=d['time_to_converged']; print({k:(v['seconds'], v['ms_per_iteration']) for k,v in t.items() if isinstance(v,dict) and 'seconds' in v})
print('ensemble', d['ensemble']['value'], 'decoder', d['decoder']['value'], 'slab outer', d['slab']['outer_iterations']['ms_per_iteration'])
lg=d['large_grid']; print({k:(round(lg[k]['value'],1), round(lg[k]['roofline']['frac'],3)) for k in ('pressure','momentum_upwind','momentum_quick')})
PY
