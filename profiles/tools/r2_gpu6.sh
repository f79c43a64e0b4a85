# round 2, GPU call 6: deferred accept side effects, k_move with 1/2 pairs per warp, the 300-locus shard
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "fast_path or pipeline or sim50 or speculation or counters" > gpurun_out/g6_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g6_tests.log
tail -4 gpurun_out/g6_tests.log
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 200 "1,1,0,1,4 1,1,0,1,2 1,1,0,1,1 2,4,0,1,4 2,4,0,1,2 2,2,0,1,4 2,6,0,1,4" > gpurun_out/g6_paths50.log 2>&1
cat gpurun_out/g6_paths50.log
IMA_BURN=300 IMA_TIMED=1 timeout 900 python profiles/tools/pipe_sweep.py sim300x256 30 "1,1,0,1,8 1,1,0,1,4 2,2,0,1,8 4,2,0,1,8" > gpurun_out/g6_sweep300.log 2>&1
cat gpurun_out/g6_sweep300.log
