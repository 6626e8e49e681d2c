mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 ) > gpurun_out/t_all.log
( timeout 200 python tests/prof_igemm_shapes.py --widths 2>&1 | grep -v "^done" ) > gpurun_out/prof_igemm_widths.txt
( timeout 200 python tests/prof_attention_shapes.py 2>&1 | grep -v "^done" ) > gpurun_out/prof_attention.txt
( timeout 700 python bench.py --shapes > gpurun_out/bench6.json ) 2> gpurun_out/bench6.err
tail -n 6 gpurun_out/t_all.log; cat gpurun_out/prof_igemm_widths.txt; cat gpurun_out/prof_attention.txt; tail -c 300 gpurun_out/bench6.err; head -c 600 gpurun_out/bench6.json
