# round 2, GPU call 17: branch-free counting sort; chain groups staggered by one proposal phase (IMA2P_STAGGER)
mkdir -p gpurun_out
rm -f gpurun_out/g17_variants.jsonl
for st in 0 1; do
  echo "stagger $st" | tee -a gpurun_out/g17_variants.jsonl
  IMA2P_STAGGER=$st IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4 3,4,0,1,4 4,4,0,1,4 2,1,0,1,4" 2>&1 | grep -v counters | tee -a gpurun_out/g17_variants.jsonl | cut -c1-200
  IMA2P_STAGGER=$st IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "2,2,0,1,8 3,2,0,1,8 4,2,0,1,8" 2>&1 | grep -v counters | tee -a gpurun_out/g17_variants.jsonl | cut -c1-200
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or pipeline or static or weights" > gpurun_out/g17_tests.log 2>&1; tail -3 gpurun_out/g17_tests.log
IMA2P_B200_LIB=$PWD/build_variants/lib_prof.so timeout 600 python profiles/tools/one_step.py sim50x128 320 3 1 4 > gpurun_out/g17_prof.log 2>&1
grep "PROFW" gpurun_out/g17_prof.log | tail -2 | sed 's/PROFW [0-9]* nev [0-9]*://g' > gpurun_out/g17_prof_split.txt
cut -c1-700 gpurun_out/g17_prof_split.txt
