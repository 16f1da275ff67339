cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tail_fused -s 3 -c 1 -o gpurun_out/c7_tail -f python tools/bench_decoder.py --batch 128 --reps 1 > gpurun_out/c7_ncu.log 2>&1; echo "ncu rc=$?"
