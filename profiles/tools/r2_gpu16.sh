# round 2, GPU call 16: counting sort (broadcast reads) in place of the bitonic network
mkdir -p gpurun_out
rm -f gpurun_out/g16_variants.jsonl
IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4 1,1,0,1,4" 2>&1 | grep -v counters | tee -a gpurun_out/g16_variants.jsonl | cut -c1-420
IMA_TIMED=1 IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "2,2,0,1,8" 2>&1 | grep -v counters | tee -a gpurun_out/g16_variants.jsonl | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or pipeline or nielsen or split or swap or capacity or speculation or hky or tupdate or static or weights" > gpurun_out/g16_tests.log 2>&1; tail -3 gpurun_out/g16_tests.log
IMA2P_B200_LIB=$PWD/build_variants/lib_prof.so timeout 600 python profiles/tools/one_step.py sim50x128 320 3 1 4 > gpurun_out/g16_prof.log 2>&1
grep "PROFT\|PROFS" gpurun_out/g16_prof.log | tail -24 > gpurun_out/g16_prof_split.txt
grep "PROFM" gpurun_out/g16_prof.log | tail -3 >> gpurun_out/g16_prof_split.txt
grep "PROFW" gpurun_out/g16_prof.log | tail -2 | sed 's/PROFW [0-9]* nev [0-9]*://g' >> gpurun_out/g16_prof_split.txt
tail -12 gpurun_out/g16_prof_split.txt | cut -c1-700
