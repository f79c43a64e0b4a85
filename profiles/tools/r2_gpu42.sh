# round 2, GPU calls 42, 44 (one GPU): the scan queue (42), the candidate loop (44) of kept rows (terms taken 32 at a time): L-mode tests, the probe at both sizes, racecheck / memcheck of the small probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lmode or l_mode or joint" > gpurun_out/g${CALL:-42}_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g${CALL:-42}_tests.log
tail -n 3 gpurun_out/g${CALL:-42}_tests.log
python profiles/tools/lmode_probe.py 1000000 512 > gpurun_out/g${CALL:-42}_probe.log 2>&1
python profiles/tools/lmode_probe.py 125000 512 >> gpurun_out/g${CALL:-42}_probe.log 2>&1
tail -n 2 gpurun_out/g${CALL:-42}_probe.log
for tool in racecheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/tools/lmode_probe.py 20011 41 all > gpurun_out/g${CALL:-42}_$tool.log 2>&1; echo "$tool rc $?"
  tail -n 1 gpurun_out/g${CALL:-42}_$tool.log | cut -c1-200
done
