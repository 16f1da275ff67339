cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_slab.py -q -x -k "two_per_pass and (world1 or 1-) or many_strips" > gpurun_out/c15_memcheck_slab.log 2>&1; echo "memcheck slab rc=$?"
tail -5 gpurun_out/c15_memcheck_slab.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sr.py -q -x -k "fused and 1" > gpurun_out/c15_memcheck_sr.log 2>&1; echo "memcheck sr rc=$?"
tail -5 gpurun_out/c15_memcheck_sr.log
