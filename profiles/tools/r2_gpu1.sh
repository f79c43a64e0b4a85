# round 2, GPU call 1: correctness of the new step structure on the device, then the timing sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or speculation or fast_path or smoke or sim50" > gpurun_out/g1_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g1_tests.log
tail -5 gpurun_out/g1_tests.log
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 200 "1,1,0,0,0 1,1,0,1,4 1,1,0,1,8 1,1,0,1,16 1,1,0,1,32" > gpurun_out/g1_paths50.log 2>&1
cat gpurun_out/g1_paths50.log
timeout 600 python profiles/tools/pipe_sweep.py sim50x128 200 "2,1,0,1,4 4,1,0,1,4 8,1,0,1,4 2,4,0,1,4 4,4,0,1,4 8,4,0,1,4 4,8,0,1,4 8,8,0,1,4 16,8,0,1,4 4,8,1,1,4 8,8,1,1,4 4,1,1,1,4 4,4,0,0,0 4,8,1,0,0" > gpurun_out/g1_sweep50.log 2>&1
cat gpurun_out/g1_sweep50.log
IMA_BURN=300 IMA_TIMED=1 timeout 900 python profiles/tools/pipe_sweep.py sim300x256 30 "1,1,0,0,0 1,1,0,1,8 1,1,0,1,32 4,4,0,1,8 8,4,0,1,8 8,8,1,1,8 4,4,0,1,32" > gpurun_out/g1_sweep300.log 2>&1
cat gpurun_out/g1_sweep300.log
IMA2P_B200_LIB=build/libima2p_b200_prof.so IMA_BURN=400 timeout 300 python profiles/tools/pipe_sweep.py sim50x128 3 "1,1,0,1,4 1,1,0,1,8" 2>&1 | grep -E "PROFW|PROFM|PROFA" | tail -300 > gpurun_out/g1_prof.log
tail -40 gpurun_out/g1_prof.log
