# round 2, GPU call 23 (2 GPUs): programmatic launches by default; two-GPU tests and the N = 2 bench
mkdir -p gpurun_out
rm -f gpurun_out/g23_variants.jsonl
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4" 2>&1 | grep -v counters | tee -a gpurun_out/g23_variants.jsonl | cut -c1-420
IMA_TIMED=1 IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "4,2,0,1,8" 2>&1 | grep -v counters | tee -a gpurun_out/g23_variants.jsonl | cut -c1-420
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "two_gpus or pipeline or fast_path or hky" > gpurun_out/g23_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g23_tests.log
tail -4 gpurun_out/g23_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 > gpurun_out/g23_bench_n2.json 2> gpurun_out/g23_bench_n2.err; echo "rc $?"; grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func" gpurun_out/g23_bench_n2.err | tail -6; cut -c1-260 gpurun_out/g23_bench_n2.json
