# round 2, GPU call 14: cycle stamps of the split-time kernels (tuning build), occupancy variants of k_split_t_fast
mkdir -p gpurun_out
IMA2P_B200_LIB=$PWD/build_variants/lib_prof.so timeout 600 python profiles/tools/one_step.py sim50x128 320 3 1 4 > gpurun_out/g14_prof.log 2>&1
grep -c PROF gpurun_out/g14_prof.log
grep "PROFT\|PROFS" gpurun_out/g14_prof.log | tail -60 > gpurun_out/g14_prof_split.txt
tail -40 gpurun_out/g14_prof_split.txt
for v in default split8 split5; do
  if [ $v = default ]; then unset IMA2P_B200_LIB; else export IMA2P_B200_LIB=$PWD/build_variants/lib_$v.so; fi
  echo "variant $v"
  IMA_TIMED=1 timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4" 2>&1 | grep -v counters | tee -a gpurun_out/g14_variants.jsonl | cut -c1-420
  IMA_TIMED=1 IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "2,2,0,1,8" 2>&1 | grep -v counters | tee -a gpurun_out/g14_variants.jsonl | cut -c1-420
done
