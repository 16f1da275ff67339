cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sr.py -q -x > gpurun_out/c20_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c20_tests.log
tail -4 gpurun_out/c20_tests.log
for w in 1 0; do
SRCFD_TC3_WIDE256=$w timeout 300 python tools/bench_decoder.py > gpurun_out/c20_dec_$w.json 2> gpurun_out/c20_dec.err; echo "dec rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c20_dec_$w.json')); print($w, {k:round(v['ms'],3) for k,v in d['paths'].items()}, d['tc_error'])"
done
