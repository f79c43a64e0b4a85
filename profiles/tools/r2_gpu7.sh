# round 2, GPU call 7 (2 GPUs): the peer-memory exchange between two processes against the one-GPU run
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 profiles/tools/sharded_check.py state_sim5_hn4 40 > gpurun_out/g7_sharded.log 2>&1; echo "rc $?" >> gpurun_out/g7_sharded.log
grep -v "^W1\|^\*\*\*\|OMP_NUM" gpurun_out/g7_sharded.log | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 profiles/tools/sharded_check.py state_sim50_hn3 40 > gpurun_out/g7_sharded50.log 2>&1; echo "rc $?" >> gpurun_out/g7_sharded50.log
grep -v "^W1\|^\*\*\*\|OMP_NUM" gpurun_out/g7_sharded50.log | tail -6
