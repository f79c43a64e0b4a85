# round 2, GPU call 43 (one GPU): the final build of the session: smoke(), the default bench line, the ABI / step-report tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c1-120
timeout 300 python -m pytest tests -m gpu -x -q -k "abi or step_report or static_eval" > gpurun_out/g43_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g43_tests.log
tail -n 2 gpurun_out/g43_tests.log
python bench.py > gpurun_out/g43_bench_n1.json 2> gpurun_out/g43_bench_n1.err; tail -c 300 gpurun_out/g43_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/g43_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['lmode']['jointp_geneval_per_sec'], d['lmode']['margincalc_geneval_per_sec'])
"
