set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -q -k "momentum" -x > gpurun_out/c3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c3_tests.log
PROBE_CHUNKS="" PROBE_PF="1,2,3,4,6,8,12" timeout 600 python tools/momentum_probe.py > gpurun_out/c3_probe.json 2> gpurun_out/c3_probe.err; echo "probe rc=$?"
tail -3 gpurun_out/c3_tests.log; cat gpurun_out/c3_probe.json
