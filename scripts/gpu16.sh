mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 600 python -m pytest tests/test_gpu_z_zdrivers.py tests/test_gpu_kernels.py tests/test_gpu_z_vae.py -q -x --timeout 300 2>&1 | tail -6 ) > gpurun_out/t_part.log
( timeout 200 python tests/prof_igemm_shapes.py 2>&1 | grep -v "^done" ) > gpurun_out/prof_igemm_now.txt
tail -n 4 gpurun_out/t_part.log; cat gpurun_out/prof_igemm_now.txt
