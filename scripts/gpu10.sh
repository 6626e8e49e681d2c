mkdir -p gpurun_out
( timeout 300 python tests/prof_igemm_roles.py --epi 2>&1 | grep -v "^done" ) > gpurun_out/prof_roles_epi.txt
cat gpurun_out/prof_roles_epi.txt
