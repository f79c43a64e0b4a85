# round 2, GPU calls 35, 36 (one GPU): warp-per-vector prefix / fold and the group skip of the scan (35), the power-of-ten table (36): L-mode tests, the probe at both sizes
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lmode or l_mode or joint" > gpurun_out/g${CALL:-35}_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g${CALL:-35}_tests.log
tail -n 3 gpurun_out/g${CALL:-35}_tests.log
python profiles/tools/lmode_probe.py 1000000 512 > gpurun_out/g${CALL:-35}_probe.log 2>&1
python profiles/tools/lmode_probe.py 125000 512 >> gpurun_out/g${CALL:-35}_probe.log 2>&1
tail -n 2 gpurun_out/g${CALL:-35}_probe.log
