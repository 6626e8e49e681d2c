mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
( timeout 120 python tests/prof_norm.py 2>&1 | grep -E "gn_fwd|done" | head -12 ) > gpurun_out/prof_norm2.txt
( timeout 600 python bench.py --shapes > gpurun_out/bench3.json ) 2> gpurun_out/bench3.err
( DDPO_BENCH_MACRO=25 timeout 400 python bench.py --phase ppo --no-cpu > gpurun_out/bench_macro25.json ) 2> gpurun_out/bench_macro25.err
timeout 400 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_sample.csv --metrics $M python bench.py --ncu sample --steps 1 --warmup 3 --no-cpu --phase sample > gpurun_out/ncu_sample.log 2>&1
timeout 500 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_train.csv --metrics $M python bench.py --ncu train --steps 1 --warmup 3 --no-cpu --phase ppo > gpurun_out/ncu_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_|layernorm_fwd" -c 10 -o gpurun_out/r2_norm_full python tests/prof_norm.py --once > gpurun_out/ncu_norm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 2 -o gpurun_out/r2_attention_full python tests/prof_attention_shapes.py --once > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm2 -c 4 -o gpurun_out/r2_igemm2_full python tests/prof_igemm_shapes.py --once > gpurun_out/ncu_igemm.log 2>&1
( timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "igemm_gn_slab or groupnorm_fwd_from_slab or layernorm_fwd or test_igemm_conv or test_attention_fwd or ddim_step_sample" 2>&1 | tail -15 ) > gpurun_out/sanitizer_memcheck.txt
cat gpurun_out/prof_norm2.txt; tail -c 600 gpurun_out/bench3.err; ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv; tail -5 gpurun_out/sanitizer_memcheck.txt
