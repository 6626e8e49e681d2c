mkdir -p gpurun_out
( timeout 300 python tests/prof_igemm_roles.py --shortk 2>&1 | grep -v "^done" ) > gpurun_out/prof_shortk.txt
cat gpurun_out/prof_shortk.txt
