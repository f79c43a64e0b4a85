# round 2, GPU call 31: ncu --set full of the hot kernels on the HBM-resident target shape (300 loci x 256 chains, one GPU's share of config 3)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_move|k_weigh|k_accept|k_split_t_fast|k_swap|k_changeu' -s 290 -c 8 -o gpurun_out/r2s7_hot300 python profiles/tools/one_step.py sim300x256 40 2 1 8 > gpurun_out/g31_ncu.log 2>&1
tail -3 gpurun_out/g31_ncu.log
ls -la gpurun_out/r2s7_hot300*
