mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gn_slab or slab_stats or layernorm_fwd or groupnorm or conv_in" 2>&1 | tail -15 ) > gpurun_out/t_k.log
( timeout 200 python tests/prof_norm.py 2>&1 | tail -40 ) > gpurun_out/prof_norm.txt
( timeout 150 python tests/prof_attention_shapes.py 2>&1 | tail -12 ) > gpurun_out/prof_attention.txt
( timeout 150 python tests/prof_igemm_shapes.py 2>&1 | tail -14 ) > gpurun_out/prof_igemm.txt
( timeout 500 python -m pytest tests/test_gpu_sd2_parity.py tests/test_gpu_z_vae.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/t_sd2.log
tail -n 4 gpurun_out/t_k.log gpurun_out/t_sd2.log
cat gpurun_out/prof_norm.txt gpurun_out/prof_attention.txt gpurun_out/prof_igemm.txt gpurun_out/sd2_parity.txt
