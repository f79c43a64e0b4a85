# round 2, GPU call 24 (8 GPUs): the bench line at N = 8 with config3 (2,048 chains x 300 loci) on the session-6 kernels
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 > gpurun_out/g24_bench_n8.json 2> gpurun_out/g24_bench_n8.err; echo "rc $?"; grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func" gpurun_out/g24_bench_n8.err | tail -6; cut -c1-260 gpurun_out/g24_bench_n8.json
