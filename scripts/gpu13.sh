mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_backward_kernels.py tests/test_gpu_kernels.py -q -k "attention" -x --timeout 300 2>&1 | tail -6 ) > gpurun_out/t_attn.log
( timeout 200 python tests/prof_attention_shapes.py 2>&1 | grep -v "^done" ) > gpurun_out/prof_attention.txt
tail -n 4 gpurun_out/t_attn.log; cat gpurun_out/prof_attention.txt
