cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for nx in 544 288 1056; do for nl in 4 3 2; do
echo "nl=$nl"; SRCFD_JTB2_FORCE=1 SRCFD_JTB2_NL=$nl timeout 300 python tools/thin_slab_probe.py $nx 4096 2>&1 | tail -1
done; done > gpurun_out/c12_thin.txt
cat gpurun_out/c12_thin.txt
