mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward_kernels.py -m gpu -q -x -k "cross_attention or attention or groupnorm_bwd or layernorm_bwd or geglu_bwd or gn_slab or groupnorm_fwd" 2>&1 | tail -25 ) > gpurun_out/t_k.log
( timeout 150 python tests/prof_attention_shapes.py 2>&1 | tail -12 ) > gpurun_out/prof_attention.txt
( timeout 120 python tests/bench_vae.py 2>&1 | tail -12 ) > gpurun_out/bench_vae.txt
( timeout 500 python -m pytest tests/test_gpu_sd2_parity.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/t_sd2.log
( timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_gpu_sd2_parity.py 2>&1 | tail -30 ) > gpurun_out/t_all.log
( timeout 700 python bench.py --shapes > gpurun_out/bench4.json ) 2> gpurun_out/bench4.err
tail -n 6 gpurun_out/t_k.log gpurun_out/t_sd2.log gpurun_out/t_all.log
cat gpurun_out/prof_attention.txt gpurun_out/bench_vae.txt; tail -c 300 gpurun_out/bench4.err; head -c 600 gpurun_out/bench4.json
