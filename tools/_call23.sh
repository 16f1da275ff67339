cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" _v2; do
  SRCFD_LIB=$GRAFT_REPO_ROOT/sr-for-cfd_b200/srcfd/_lib/libsrcfd$v.so PROBE_CHUNKS="" PROBE_PF="" timeout 600 python tools/momentum_probe.py > gpurun_out/c23_probe$v.json 2> gpurun_out/c23_probe$v.err; echo "probe$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/c23_probe$v.json')); print({k:v['glups'] if isinstance(v,dict) else v for k,v in d.items()})"
done
