mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward_kernels.py -q -x -k "igemm or conv or linear or wgrad" --timeout 300 2>&1 | tail -3 ) > gpurun_out/t_gemm.log
( timeout 200 python tests/prof_igemm_shapes.py 2>&1 | grep -v "^done" ) > gpurun_out/prof_igemm_now.txt
( timeout 200 python tests/prof_wgrad.py 2>&1 | tail -3 ) > gpurun_out/prof_wgrad.txt
tail -n 2 gpurun_out/t_gemm.log; cat gpurun_out/prof_igemm_now.txt gpurun_out/prof_wgrad.txt
