# round 2, GPU call 41 (eight GPUs): the N = 8 bench line with this session's L-mode kernels (config 3 and the CPU baseline were measured with
# the same M-mode kernels in call 24, profiles/r2s6_bench_n8.json, and are left out here to keep the eight-GPU box time short)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --no-cpu-baseline --no-config3 > gpurun_out/g41_bench_n8.json 2> gpurun_out/g41_bench_n8.err; echo "rc $?"
grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func" gpurun_out/g41_bench_n8.err | tail -n 5
python -c "
import json
d=json.loads(open('gpurun_out/g41_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['lmode'])
"
