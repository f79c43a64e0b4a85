# round 2, GPU call 8: the bench line at N = 1 (synthetic and real input), reference arm
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/g8_bench_n1.json 2> gpurun_out/g8_bench_n1.err; echo "rc $?"; tail -3 gpurun_out/g8_bench_n1.err; cut -c1-1500 gpurun_out/g8_bench_n1.json
timeout 900 python bench.py --data synthetic --no-lmode --no-models > gpurun_out/g8_bench_n1_real.json 2> gpurun_out/g8_bench_n1_real.err; echo "rc $?"; tail -3 gpurun_out/g8_bench_n1_real.err; cut -c1-700 gpurun_out/g8_bench_n1_real.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/g8_bench_ref.json 2>/dev/null; cut -c1-400 gpurun_out/g8_bench_ref.json
