cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PROBE_SFN=1 PROBE_CHUNKS="" PROBE_PF="" timeout 600 python tools/momentum_probe.py > gpurun_out/c11_probe.json 2> gpurun_out/c11_probe.err; echo "probe rc=$?"
cat gpurun_out/c11_probe.json
