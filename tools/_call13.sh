cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slab_sweep2 -s 4 -c 1 -o gpurun_out/c13_sweep2_up -f python tools/sweep2_once.py > gpurun_out/c13_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/bench_decoder.py --batch 256 --reps 1 > /dev/null 2>&1; echo "dec plain rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c13_dec_launches.csv python tools/bench_decoder.py --batch 256 --reps 1 > gpurun_out/c13_dec_ncu.log 2>&1; echo "ncu rc=$?"
