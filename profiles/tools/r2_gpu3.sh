# round 2, GPU call 3: speculation depth x chain groups x graph depth on the one-barrier accept sweep
mkdir -p gpurun_out
timeout 600 python profiles/tools/pipe_sweep.py sim50x128 200 "1,1,0,1,4,3 1,1,0,1,4,2 1,1,0,1,4,4 2,4,0,1,4,2 2,4,0,1,4,3 2,4,0,1,4,4 4,4,0,1,4,2 4,4,0,1,4,3 2,8,0,1,4,2 4,8,0,1,4,2 2,4,1,1,4,2 4,4,1,1,4,2 3,4,0,1,4,2 2,2,0,1,4,2" > gpurun_out/g3_spec50.log 2>&1
cat gpurun_out/g3_spec50.log
