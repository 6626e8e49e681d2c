mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "igemm or conv or linear" -x --timeout 300 2>&1 | tail -4 ) > gpurun_out/t_igemm.log
( DDPO_PROF_ROOT=tests/microbench/old timeout 120 python tests/prof_igemm_roles.py --ab 2>&1 | grep -v "^done" ) > gpurun_out/prof_roles_old.txt
( timeout 300 python tests/prof_igemm_roles.py 2>&1 | grep -v "^done" ) > gpurun_out/prof_roles.txt
( timeout 300 python tests/prof_igemm_shapes.py --widths 2>&1 | grep -v "^done" ) > gpurun_out/prof_igemm_widths.txt
tail -n 3 gpurun_out/t_igemm.log; echo OLD; cat gpurun_out/prof_roles_old.txt; echo NEW; cat gpurun_out/prof_roles.txt; cat gpurun_out/prof_igemm_widths.txt
