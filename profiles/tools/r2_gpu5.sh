# round 2, GPU call 5: accept sweep with batched transcendentals / log-uniform decisions, staged accept_t
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "static or proposals or fast_path or pipeline or sim50 or split_time or gamma or numerics or speculation" > gpurun_out/g5_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g5_tests.log
tail -4 gpurun_out/g5_tests.log
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 200 "1,1,0,1,4 2,4,0,1,4 1,1,0,1,4,4 1,1,0,1,4,2" > gpurun_out/g5_paths50.log 2>&1
cat gpurun_out/g5_paths50.log
IMA2P_B200_LIB=build/libima2p_b200_prof.so IMA_BURN=400 timeout 300 python profiles/tools/pipe_sweep.py sim50x128 2 "1,1,0,1,4" 2>&1 | grep -E "PROFA" | tail -30 > gpurun_out/g5_prof.log
cat gpurun_out/g5_prof.log
