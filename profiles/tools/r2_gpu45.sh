# round 2, GPU call 45 (one GPU): ncu --set full of the final joint kernels (one 512-vector pass over 1,000,000 rows)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_joint_terms|k_joint_scan|k_joint_prefix|k_joint_fold' -s 7 -c 7 -o gpurun_out/r2s8_lmode_final python profiles/tools/lmode_probe.py 1000000 512 > gpurun_out/g45_ncu.log 2>&1
tail -n 2 gpurun_out/g45_ncu.log | cut -c1-200
ls -la gpurun_out/r2s8_lmode_final*
