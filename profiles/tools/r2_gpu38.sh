# round 2, GPU call 38 (two GPUs): the two-GPU tests and the N = 2 bench line (config 3 and the CPU baseline left to the driver's run)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "two_gpus" > gpurun_out/g38_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g38_tests.log
tail -n 3 gpurun_out/g38_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --no-cpu-baseline --no-config3 > gpurun_out/g38_bench_n2.json 2> gpurun_out/g38_bench_n2.err; echo "rc $?"
grep -v "^W1\|^\*\*\*\|OMP_NUM\|UserWarning\|return func" gpurun_out/g38_bench_n2.err | tail -n 5
python -c "
import json
d=json.loads(open('gpurun_out/g38_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['lmode'])
"
