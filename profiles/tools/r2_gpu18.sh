# round 2, GPU call 18: continued fraction by its convergents (no scan, no division per term), exponentials only when the fallback needs them
mkdir -p gpurun_out
rm -f gpurun_out/g18_variants.jsonl
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4 1,1,0,1,4 2,4,0,1,4,2 2,4,0,1,4,4" 2>&1 | grep -v counters | tee -a gpurun_out/g18_variants.jsonl | cut -c1-420
IMA_TIMED=1 IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "2,2,0,1,8 4,2,0,1,8" 2>&1 | grep -v counters | tee -a gpurun_out/g18_variants.jsonl | cut -c1-420
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not two_gpus and not front_end" > gpurun_out/g18_tests.log 2>&1; tail -3 gpurun_out/g18_tests.log
IMA2P_B200_LIB=$PWD/build_variants/lib_prof.so timeout 600 python profiles/tools/one_step.py sim50x128 320 3 1 4 > gpurun_out/g18_prof.log 2>&1
grep "PROFA" gpurun_out/g18_prof.log | tail -15 > gpurun_out/g18_prof_accept.txt
grep "PROFT" gpurun_out/g18_prof.log | tail -4 >> gpurun_out/g18_prof_accept.txt
cat gpurun_out/g18_prof_accept.txt
