# round 2, GPU call 13: the whole GPU suite on the final kernels, the bench line at N = 1, launch list + ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/g13_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g13_tests.log
tail -4 gpurun_out/g13_tests.log
timeout 900 python bench.py > gpurun_out/g13_bench_n1.json 2> gpurun_out/g13_bench_n1.err; echo "rc $?"; tail -2 gpurun_out/g13_bench_n1.err; cut -c1-200 gpurun_out/g13_bench_n1.json
timeout 600 python bench.py --data synthetic --no-lmode --no-models --no-cpu-baseline > gpurun_out/g13_bench_n1_synth.json 2>/dev/null; cut -c1-200 gpurun_out/g13_bench_n1_synth.json
# every launch of a few steps with its device time (cold cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 60 --csv --log-file gpurun_out/r2s5_launches_sim50x128.csv python profiles/tools/one_step.py sim50x128 300 6 1 4 > gpurun_out/g13_launches.log 2>&1
# the hot kernels, one launch each, full sections (the graph burn-in launches 9 kernels a step; the eager steps follow)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_move|k_weigh|k_accept|k_split_t_fast|k_swap|k_changeu' -s 2108 -c 7 -o gpurun_out/r2s5_hot python profiles/tools/one_step.py sim50x128 300 3 1 4 > gpurun_out/g13_ncu.log 2>&1
tail -3 gpurun_out/g13_ncu.log
ls -la gpurun_out/r2s5*
