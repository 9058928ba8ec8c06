mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_tc_kernel<64, 3" -s 2 -c 1 -f -o gpurun_out/prof_dsum python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/ncu_dsum.log 2>&1
tail -3 gpurun_out/ncu_dsum.log
