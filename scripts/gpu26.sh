mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
( timeout 150 $N -k regex:wgrad2_kernel -c 3 -f -o gpurun_out/r2e_wgrad2 python tests/prof_wgrad.py ) > gpurun_out/ncu_wgrad.log 2>&1
( timeout 150 $N -k regex:igemm2 -c 4 -f -o gpurun_out/r2e_igemm2 python tests/prof_igemm_shapes.py --once ) > gpurun_out/ncu_igemm2.log 2>&1
ls -la gpurun_out/r2e_*.ncu-rep
