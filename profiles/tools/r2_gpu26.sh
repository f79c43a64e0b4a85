# round 2, GPU call 26: compute-sanitizer over a few whole steps of small engines (infinite sites, HKY, stepwise)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/tools/sanitize.py 6 > gpurun_out/g26_$tool.log 2>&1; echo "$tool rc $?"
  grep -c "ERROR\|Error\|error" gpurun_out/g26_$tool.log; tail -4 gpurun_out/g26_$tool.log | cut -c1-200
done
