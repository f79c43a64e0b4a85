# round 2, GPU call 19: the whole GPU suite on the session-6 kernels, the bench line at N = 1 (real and synthetic input), launch list + ncu --set full
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/g19_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g19_tests.log
tail -4 gpurun_out/g19_tests.log
timeout 900 python bench.py > gpurun_out/g19_bench_n1.json 2> gpurun_out/g19_bench_n1.err; echo "rc $?"; tail -2 gpurun_out/g19_bench_n1.err; cut -c1-200 gpurun_out/g19_bench_n1.json
timeout 600 python bench.py --data synthetic --no-lmode --no-models --no-cpu-baseline > gpurun_out/g19_bench_n1_synth.json 2>/dev/null; cut -c1-200 gpurun_out/g19_bench_n1_synth.json
timeout 600 python bench.py --impl reference > gpurun_out/g19_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/g19_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5100 -c 70 --csv --log-file gpurun_out/r2s6_launches_sim50x128.csv python profiles/tools/one_step.py sim50x128 300 4 1 4 > gpurun_out/g19_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_move|k_weigh|k_accept|k_split_t_fast|k_swap|k_changeu' -s 4208 -c 7 -o gpurun_out/r2s6_hot python profiles/tools/one_step.py sim50x128 300 3 1 4 > gpurun_out/g19_ncu.log 2>&1
tail -3 gpurun_out/g19_ncu.log
ls -la gpurun_out/r2s6*
