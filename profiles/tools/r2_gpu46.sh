# round 2, GPU call 46 (one GPU): the scan loads four groups of rows ahead: L-mode tests and the probe
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "lmode or joint" > gpurun_out/g46_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g46_tests.log
tail -n 2 gpurun_out/g46_tests.log
python profiles/tools/lmode_probe.py 1000000 512 > gpurun_out/g46_probe.log 2>&1
python profiles/tools/lmode_probe.py 125000 512 >> gpurun_out/g46_probe.log 2>&1
tail -n 2 gpurun_out/g46_probe.log
