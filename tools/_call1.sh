set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_slab.py -q -k "momentum" -x > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c1_tests.log
timeout 600 python tools/momentum_probe.py > gpurun_out/c1_probe.json 2> gpurun_out/c1_probe.err; echo "probe rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slab_sweep2 -s 4 -c 1 -o gpurun_out/c1_sweep2_up -f python tools/sweep2_once.py > gpurun_out/c1_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/c1_tests.log; cat gpurun_out/c1_probe.json
