cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PROBE_CHUNKS="" PROBE_PF="" timeout 600 python tools/momentum_probe.py > gpurun_out/c22_probe.json 2> gpurun_out/c22_probe.err; echo "probe rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c22_probe.json')); print({k:v['glups'] if isinstance(v,dict) else v for k,v in d.items()})"
timeout 600 python -m pytest tests/test_gpu_slab.py -q -x -k "momentum or outer" > gpurun_out/c22_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c22_tests.log
tail -3 gpurun_out/c22_tests.log
