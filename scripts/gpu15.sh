mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 ) > gpurun_out/t_all.log
( timeout 700 python bench.py --shapes > gpurun_out/bench7.json ) 2> gpurun_out/bench7.err
tail -n 3 gpurun_out/t_all.log; tail -c 200 gpurun_out/bench7.err; head -c 400 gpurun_out/bench7.json
