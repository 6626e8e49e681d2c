mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gn_slab or slab_stats or layernorm_fwd or groupnorm or conv_in or attention_fwd" 2>&1 | tail -25 ) > gpurun_out/t_k.log
( timeout 200 python tests/prof_norm.py 2>&1 | tail -30 ) > gpurun_out/prof_norm.txt
( timeout 600 python -m pytest tests/test_gpu_sd2_parity.py -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/t_sd2.log
( timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_sd2_parity.py 2>&1 | tail -40 ) > gpurun_out/t_all.log
( timeout 700 python bench.py --shapes > gpurun_out/bench2.json ) 2> gpurun_out/bench2.err
tail -n 3 gpurun_out/t_k.log gpurun_out/t_sd2.log gpurun_out/t_all.log
cat gpurun_out/prof_norm.txt
