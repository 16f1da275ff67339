cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/c24_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c24_tests.log
tail -3 gpurun_out/c24_tests.log
timeout 600 python bench.py --steps 100 --no-extras > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c24_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['step_breakdown_ms'])
PY
