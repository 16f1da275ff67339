cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sr.py -q -x  > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c9_tests.log
tail -4 gpurun_out/c9_tests.log
timeout 600 python tools/bench_decoder.py > gpurun_out/c9_dec.json 2> gpurun_out/c9_dec.err; echo "dec rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c9_dec.json')); print({k:round(v['ms'],3) for k,v in d['paths'].items()}, d['tc_error'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tail_fused -s 3 -c 1 -o gpurun_out/c9_tail -f python tools/bench_decoder.py --batch 128 --reps 1 > gpurun_out/c9_ncu.log 2>&1; echo "ncu rc=$?"
