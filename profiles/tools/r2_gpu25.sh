# round 2, GPU call 25: the whole GPU suite on the final build, the bench lines at N = 1 (real and synthetic input), the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/g25_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g25_tests.log
tail -4 gpurun_out/g25_tests.log
timeout 900 python bench.py > gpurun_out/g25_bench_n1.json 2> gpurun_out/g25_bench_n1.err; echo "rc $?"; tail -2 gpurun_out/g25_bench_n1.err; cut -c1-200 gpurun_out/g25_bench_n1.json
timeout 600 python bench.py --data synthetic --no-lmode --no-models --no-cpu-baseline > gpurun_out/g25_bench_n1_synth.json 2>/dev/null; cut -c1-200 gpurun_out/g25_bench_n1_synth.json
timeout 600 python bench.py --impl reference > gpurun_out/g25_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/g25_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
