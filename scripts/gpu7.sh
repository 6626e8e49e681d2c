mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 120 tests/microbench/l2_tma_bw ) > gpurun_out/l2_tma_bw.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "igemm or conv or linear" -x --timeout 300 2>&1 | tail -8 ) > gpurun_out/t_igemm.log
( timeout 300 python tests/prof_igemm_shapes.py --widths 2>&1 | grep -v "^done" ) > gpurun_out/prof_igemm_widths.txt
( timeout 600 python -m pytest tests/test_gpu_z_vae.py tests/test_gpu_training.py -q --timeout 400 2>&1 | tail -12 ) > gpurun_out/t_vae_train.log
cat gpurun_out/l2_tma_bw.txt; tail -n 5 gpurun_out/t_igemm.log; cat gpurun_out/prof_igemm_widths.txt; tail -n 8 gpurun_out/t_vae_train.log; cat gpurun_out/ppo_grad_parity_TINY.txt
