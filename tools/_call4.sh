cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" _v2 _v3; do
  SRCFD_LIB=$GRAFT_REPO_ROOT/sr-for-cfd_b200/srcfd/_lib/libsrcfd$v.so PROBE_CHUNKS="" PROBE_PF="" timeout 600 python tools/momentum_probe.py > gpurun_out/c4_probe$v.json 2> gpurun_out/c4_probe$v.err; echo "probe$v rc=$?"
  cat gpurun_out/c4_probe$v.json
done
