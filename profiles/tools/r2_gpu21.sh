# round 2, GPU call 21: programmatic dependent launch inside the step graph (IMA2P_PDL), A/B
mkdir -p gpurun_out
rm -f gpurun_out/g21_variants.jsonl
for pdl in 0 1; do
  echo "pdl $pdl" | tee -a gpurun_out/g21_variants.jsonl
  IMA2P_PDL=$pdl timeout 600 python profiles/tools/pipe_sweep.py sim50x128 400 "2,4,0,1,4 1,1,0,1,4 2,1,0,1,4" 2>&1 | tee -a gpurun_out/g21_variants.jsonl | cut -c1-200
  IMA2P_PDL=$pdl IMA_BURN=300 timeout 600 python profiles/tools/pipe_sweep.py sim300x256 60 "4,2,0,1,8" 2>&1 | tee -a gpurun_out/g21_variants.jsonl | cut -c1-200
done
IMA2P_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or fast or speculation or swap or long_run" > gpurun_out/g21_tests.log 2>&1; tail -3 gpurun_out/g21_tests.log
