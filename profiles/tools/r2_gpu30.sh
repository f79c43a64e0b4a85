# round 2, GPU call 30: log h as two logarithms beside log x (no division after the convergents)
mkdir -p gpurun_out
rm -f gpurun_out/g30_variants.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "speculation or accept or static or gamma or numerics or long_run or pipeline" > gpurun_out/g30_tests.log 2>&1; tail -3 gpurun_out/g30_tests.log
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4" 2>&1 | grep -v counters | tee -a gpurun_out/g30_variants.jsonl | cut -c1-330
IMA_TIMED=1 IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "4,2,0,1,8" 2>&1 | grep -v counters | tee -a gpurun_out/g30_variants.jsonl | cut -c1-330
