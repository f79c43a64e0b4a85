# round 2, GPU call 9 (2 GPUs): the bench line at N = 2 with config3, and the reference arm beside it
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 > gpurun_out/g9_bench_n2.json 2> gpurun_out/g9_bench_n2.err; echo "rc $?"; grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func" gpurun_out/g9_bench_n2.err | tail -8; cut -c1-300 gpurun_out/g9_bench_n2.json
