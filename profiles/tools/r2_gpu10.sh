# round 2, GPU call 10: HKY partial recompute on the device, the whole GPU suite, the bench line with the per-model section
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/g10_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g10_tests.log
tail -5 gpurun_out/g10_tests.log
timeout 900 python bench.py > gpurun_out/g10_bench_n1.json 2> gpurun_out/g10_bench_n1.err; echo "rc $?"; tail -3 gpurun_out/g10_bench_n1.err; cut -c1-300 gpurun_out/g10_bench_n1.json
