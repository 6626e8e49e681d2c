mkdir -p gpurun_out
( timeout 120 tests/microbench/l2_tma_bw ) > gpurun_out/l2_tma_bw.txt 2>&1
( DDPO_PROF_ROOT=tests/microbench/old timeout 120 python tests/prof_igemm_roles.py --ab 2>&1 | grep -v "^done" ) > gpurun_out/prof_roles_old.txt
( timeout 120 python tests/prof_igemm_roles.py --ab 2>&1 | grep -v "^done" ) > gpurun_out/prof_roles_new_ab.txt
( timeout 300 python tests/prof_igemm_roles.py 2>&1 | grep -v "^done" ) > gpurun_out/prof_roles.txt
( timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "igemm" -x --timeout 300 2>&1 | tail -4 ) > gpurun_out/t_igemm.log
cat gpurun_out/l2_tma_bw.txt; echo OLD; cat gpurun_out/prof_roles_old.txt; echo NEW; cat gpurun_out/prof_roles_new_ab.txt; cat gpurun_out/prof_roles.txt; tail -n 3 gpurun_out/t_igemm.log
