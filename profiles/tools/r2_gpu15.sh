# round 2, GPU call 15: smaller code (rolled sort, one body for both split-time updates), batched staging loads, k_accept_t in one trip
mkdir -p gpurun_out
rm -f gpurun_out/g15_variants.jsonl
for v in default split5; do
  if [ $v = default ]; then unset IMA2P_B200_LIB; else export IMA2P_B200_LIB=$PWD/build_variants/lib_$v.so; fi
  echo "variant $v"
  IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4 1,1,0,1,4" 2>&1 | grep -v counters | tee -a gpurun_out/g15_variants.jsonl | cut -c1-420
  IMA_TIMED=1 IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "2,2,0,1,8" 2>&1 | grep -v counters | tee -a gpurun_out/g15_variants.jsonl | cut -c1-420
done
unset IMA2P_B200_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or pipeline or nielsen or split or swap or capacity or speculation or hky or tupdate" > gpurun_out/g15_tests.log 2>&1; tail -3 gpurun_out/g15_tests.log
IMA2P_B200_LIB=$PWD/build_variants/lib_prof.so timeout 600 python profiles/tools/one_step.py sim50x128 320 3 1 4 > gpurun_out/g15_prof.log 2>&1
grep "PROFT\|PROFS" gpurun_out/g15_prof.log | tail -40 > gpurun_out/g15_prof_split.txt
grep "PROFM" gpurun_out/g15_prof.log | tail -5 >> gpurun_out/g15_prof_split.txt
grep "PROFW" gpurun_out/g15_prof.log | tail -3 | cut -c1-900 >> gpurun_out/g15_prof_split.txt
tail -30 gpurun_out/g15_prof_split.txt | cut -c1-300
