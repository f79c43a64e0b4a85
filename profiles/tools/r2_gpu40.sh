# round 2, GPU call 40 (one GPU): compute-sanitizer over the L-mode kernels of this session (k_joint_terms with a ragged last tile and chunk,
# k_joint_prefix / _scan / _fold, every joint model type, k_marginal_many / k_reduce_many) and the two-slot step report
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/tools/lmode_probe.py 20011 41 all > gpurun_out/g40_$tool.log 2>&1; echo "$tool rc $?"
  tail -n 3 gpurun_out/g40_$tool.log | cut -c1-300
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "step_report" > gpurun_out/g40_memcheck_report.log 2>&1; echo "report rc $?"
tail -n 3 gpurun_out/g40_memcheck_report.log | cut -c1-300
