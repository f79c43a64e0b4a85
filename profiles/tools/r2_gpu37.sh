# round 2, GPU call 37 (one GPU): compact coefficient upload of jointp; L-mode tests, the probe, smoke() and the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lmode or l_mode or joint or abi or step_report" > gpurun_out/g37_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g37_tests.log
tail -n 3 gpurun_out/g37_tests.log
python profiles/tools/lmode_probe.py 1000000 512 > gpurun_out/g37_probe.log 2>&1
python profiles/tools/lmode_probe.py 125000 512 >> gpurun_out/g37_probe.log 2>&1
tail -n 2 gpurun_out/g37_probe.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c1-200
python bench.py > gpurun_out/g37_bench_n1.json 2> gpurun_out/g37_bench_n1.err; tail -c 300 gpurun_out/g37_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/g37_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['lmode']['jointp_geneval_per_sec'], d['lmode']['margincalc_geneval_per_sec'])
"
