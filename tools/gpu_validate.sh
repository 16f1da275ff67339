#!/bin/bash
# One GPU-box validation pass: the GPU parity suite, the driver's smoke(), and the default bench line.
#   gpurun --timeout 1200 -- 'bash tools/gpu_validate.sh'
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/validate_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/validate_tests.log
tail -4 gpurun_out/validate_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/validate_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/validate_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['step_breakdown_ms'])
lg=d['large_grid']; print({k:(round(lg[k]['value'],1), round(lg[k]['roofline']['frac'],3)) for k in ('pressure','momentum_upwind','momentum_quick')})
sl=d['slab']; print('slab outer', sl['outer_iterations']['ms_per_iteration'], 'developed', sl.get('outer_iterations_developed',{}).get('ms_per_iteration'))
t=d.get('time_to_converged',{}); print({k:v.get('seconds') for k,v in t.items() if isinstance(v,dict) and 'seconds' in v})
print('ensemble', d['ensemble']['value'], 'decoder', d['decoder']['value'], {k:round(v['ms'],3) for k,v in d['decoder']['paths'].items()})
PY
