mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
( timeout 300 $N -k regex:wgrad2_kernel -c 3 -f -o gpurun_out/r2d_wgrad2 python tests/prof_wgrad.py --full ) > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out/r2d_wgrad2.ncu-rep; grep -i "wgrad" gpurun_out/ncu_wgrad.log | head -8
