# round 2, GPU call 20: uploads overlapping the steps (e2e), the levelled mutation-scalar walk (HKY / stepwise), launch list + ncu --set full
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hky or scalar or changeu or uupdate or packed or stepwise or joint or long_run or full_schedule or models" > gpurun_out/g20_tests.log 2>&1; tail -3 gpurun_out/g20_tests.log
timeout 900 python bench.py --no-lmode > gpurun_out/g20_bench_n1.json 2> gpurun_out/g20_bench_n1.err; echo "rc $?"; tail -2 gpurun_out/g20_bench_n1.err; cut -c1-200 gpurun_out/g20_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 80 --csv --log-file gpurun_out/r2s6_launches_sim50x128.csv python profiles/tools/one_step.py sim50x128 300 4 1 4 > gpurun_out/g20_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_move|k_weigh|k_accept|k_split_t_fast|k_swap|k_changeu' -s 2108 -c 7 -o gpurun_out/r2s6_hot python profiles/tools/one_step.py sim50x128 300 3 1 4 > gpurun_out/g20_ncu.log 2>&1
tail -3 gpurun_out/g20_ncu.log
ls -la gpurun_out/r2s6*
