set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -q -k "momentum" -x > gpurun_out/c2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c2_tests.log
PROBE_CHUNKS="" timeout 600 python tools/momentum_probe.py > gpurun_out/c2_probe.json 2> gpurun_out/c2_probe.err; echo "probe rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slab_sweep2 -s 4 -c 1 -o gpurun_out/c2_sweep2_up -f python tools/sweep2_once.py > gpurun_out/c2_ncu.log 2>&1; echo "ncu rc=$?"
PROBE_QUICK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slab_sweep2 -s 4 -c 1 -o gpurun_out/c2_sweep2_quick -f python tools/sweep2_once.py > gpurun_out/c2_ncu_q.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/c2_tests.log; cat gpurun_out/c2_probe.json
