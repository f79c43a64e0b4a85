# round 2, GPU call 2: fast-path identity on the device + stage cycles after the first optimisation pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or fast_path or sim50" > gpurun_out/g2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g2_tests.log
tail -5 gpurun_out/g2_tests.log
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 200 "1,1,0,0,0 1,1,0,1,4 1,1,0,1,8 2,4,0,1,4" > gpurun_out/g2_paths50.log 2>&1
cat gpurun_out/g2_paths50.log
IMA_BURN=300 IMA_TIMED=1 timeout 900 python profiles/tools/pipe_sweep.py sim300x256 30 "1,1,0,1,8 1,1,0,1,16 2,2,0,1,8" > gpurun_out/g2_sweep300.log 2>&1
cat gpurun_out/g2_sweep300.log
IMA2P_B200_LIB=build/libima2p_b200_prof.so IMA_BURN=400 timeout 300 python profiles/tools/pipe_sweep.py sim50x128 3 "1,1,0,1,4" 2>&1 | grep -E "PROFW|PROFM|PROFA" | tail -60 > gpurun_out/g2_prof.log
tail -30 gpurun_out/g2_prof.log
