cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sr.py -q -x > gpurun_out/c6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c6_tests.log
tail -15 gpurun_out/c6_tests.log
timeout 600 python tools/bench_decoder.py > gpurun_out/c6_dec.json 2> gpurun_out/c6_dec.err; echo "dec rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c6_dec.json')); print({k:round(v['ms'],3) for k,v in d['paths'].items()}, d['tc_error'])"
