mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 200 python tests/prof_norm.py 2>&1 | grep -E "ln_bwd|gn_bwd|GROUP" | head -12 ) > gpurun_out/prof_norm3.txt
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/t_all.log
( timeout 700 python bench.py --shapes > gpurun_out/bench5.json ) 2> gpurun_out/bench5.err
( timeout 400 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench5_ref.json ) 2> gpurun_out/bench5_ref.err
tail -n 6 gpurun_out/t_all.log; cat gpurun_out/prof_norm3.txt; tail -c 300 gpurun_out/bench5.err; head -c 500 gpurun_out/bench5.json; echo; head -c 700 gpurun_out/bench5_ref.json
