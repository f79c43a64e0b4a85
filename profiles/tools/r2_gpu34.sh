# round 2, GPU call 34 (one GPU): L-mode probe timed, then ncu --set full of the joint / lock-step kernels (125,000 rows = one GPU's share of config 4 at N = 8, and 1,000,000)
mkdir -p gpurun_out
python profiles/tools/lmode_probe.py 1000000 512 > gpurun_out/g34_probe.log 2>&1
python profiles/tools/lmode_probe.py 125000 512 >> gpurun_out/g34_probe.log 2>&1
tail -n 3 gpurun_out/g34_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_joint_terms|k_joint_scan|k_joint_prefix|k_joint_fold|k_marginal_many' -c 10 -o gpurun_out/r2s8_lmode python profiles/tools/lmode_probe.py 1000000 512 > gpurun_out/g34_ncu.log 2>&1
tail -n 3 gpurun_out/g34_ncu.log
ls -la gpurun_out/r2s8_lmode*
