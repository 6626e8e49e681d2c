mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
( timeout 120 python tests/prof_attention_shapes.py 2>&1 | tail -12 ) > gpurun_out/prof_attention.txt
( timeout 120 python tests/prof_igemm_shapes.py 2>&1 | tail -14 ) > gpurun_out/prof_igemm.txt
timeout 400 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_sample.csv --metrics $M python bench.py --ncu sample --steps 1 --warmup 3 --no-cpu --phase sample > gpurun_out/ncu_sample.log 2>&1
timeout 500 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_train.csv --metrics $M python bench.py --ncu train --steps 1 --warmup 3 --no-cpu --phase ppo > gpurun_out/ncu_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_|layernorm_fwd" -c 12 -o gpurun_out/r2_norm_full python tests/prof_norm.py --once > gpurun_out/ncu_norm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -c 2 -o gpurun_out/r2_attention_full python tests/prof_attention_shapes.py --once > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm2 -c 4 -o gpurun_out/r2_igemm2_full python tests/prof_igemm_shapes.py --once > gpurun_out/ncu_igemm.log 2>&1
cat gpurun_out/prof_attention.txt gpurun_out/prof_igemm.txt
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
