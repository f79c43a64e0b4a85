# round 2, GPU call 11 (2 GPUs): the two-rank front end on two GPUs, the tests added since call 10, the N = 2 bench with the device-collective jointp
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "two_gpus or capacity or hky_partials or report_head or fast_path" > gpurun_out/g11_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g11_tests.log
tail -5 gpurun_out/g11_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 > gpurun_out/g11_bench_n2.json 2> gpurun_out/g11_bench_n2.err; echo "rc $?"; grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func" gpurun_out/g11_bench_n2.err | tail -6; cut -c1-260 gpurun_out/g11_bench_n2.json
