mkdir -p gpurun_out
( timeout 200 python tests/prof_igemm_shapes.py --once --lowres 2>&1 | grep "pair=[12]" ) > gpurun_out/prof_lowres.txt
( timeout 200 python tests/prof_norm.py 2>&1 | grep -E "gn_bwd|ln_bwd" | head -8 ) > gpurun_out/prof_norm4.txt
( timeout 300 python -m pytest tests/test_gpu_backward_kernels.py -q -k "groupnorm or gn" --timeout 300 2>&1 | tail -3 ) > gpurun_out/t_gn.log
cat gpurun_out/prof_lowres.txt gpurun_out/prof_norm4.txt; tail -n 2 gpurun_out/t_gn.log
