cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for g in 144 32 16 8; do
SRCFD_TC3_L1_CTAS=$g timeout 600 python tools/bench_decoder.py > gpurun_out/c10_dec_$g.json 2> gpurun_out/c10_dec.err; echo "dec rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c10_dec_$g.json')); print($g, {k:round(v['ms'],3) for k,v in d['paths'].items()}, d['tc_error'])"
done
timeout 900 python -m pytest tests/test_gpu_sr.py -q -x > gpurun_out/c10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c10_tests.log
tail -3 gpurun_out/c10_tests.log
