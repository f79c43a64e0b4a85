# round 2, GPU call 29: cycle stamps of the accept sweep, a warp per locus (depth 8) against a warp per term (depth 3)
mkdir -p gpurun_out
for sp in 3 8; do
  IMA_SPEC=$sp IMA2P_B200_LIB=$PWD/build_variants/lib_prof.so timeout 600 python profiles/tools/one_step.py sim50x128 320 3 1 4 > gpurun_out/g29_prof_$sp.log 2>&1
  echo "spec $sp"; grep "PROFA" gpurun_out/g29_prof_$sp.log | tail -8
done
