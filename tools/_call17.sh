cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_convT2x2_tc3|k_convT3x3_l1|k_col2im" -s 8 -c 6 -o gpurun_out/c17_layers -f python tools/bench_decoder.py --batch 128 --reps 1 > gpurun_out/c17_ncu.log 2>&1; echo "ncu rc=$?"
