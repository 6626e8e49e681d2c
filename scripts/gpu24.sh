mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 ) > gpurun_out/t_all.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 700 python bench.py --shapes > gpurun_out/bench8.json ) 2> gpurun_out/bench8.err
( timeout 400 python bench.py --steps 10 --warmup 3 --scaling strong --no-cpu > gpurun_out/bench_n1_strong.json ) 2> gpurun_out/bench_n1_strong.err
tail -n 3 gpurun_out/t_all.log; cat gpurun_out/smoke.log; head -c 300 gpurun_out/bench8.json; echo; head -c 200 gpurun_out/bench_n1_strong.json
