cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --no-ttc > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c25_bench.json') if l.startswith('{')][-1])
print(d['slab']['outer_iterations'])
print(d['slab'].get('outer_iterations_developed'))
PY
tail -3 gpurun_out/c25_bench.err
