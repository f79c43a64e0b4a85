# round 2, GPU call 4: parity on the device with the register sort, timings, and an ncu --set full capture of the hot kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "static or proposals or fast_path or pipeline or sim50 or split_time or nielsen" > gpurun_out/g4_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g4_tests.log
tail -4 gpurun_out/g4_tests.log
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 200 "1,1,0,1,4 2,4,0,1,4" > gpurun_out/g4_paths50.log 2>&1
cat gpurun_out/g4_paths50.log
# burn 300 graph steps = 5 matched kernels each; capture the 5 hot kernels of the second eager step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_move|k_weigh|k_accept|k_split_t_fast' -s 1505 -c 5 -o gpurun_out/r2s2_hot python profiles/tools/one_step.py sim50x128 300 3 1 4 > gpurun_out/g4_ncu.log 2>&1
tail -5 gpurun_out/g4_ncu.log
ls -la gpurun_out/*.ncu-rep
