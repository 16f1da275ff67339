cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/c5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c5_tests.log
tail -6 gpurun_out/c5_tests.log
( time timeout 900 python bench.py ) > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/c5_bench.err
