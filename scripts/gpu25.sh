mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_kernels.py -q -k "groupnorm or gn or layernorm" --timeout 300 2>&1 | tail -3 ) > gpurun_out/t_gn.log
tail -n 2 gpurun_out/t_gn.log
