mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_sample.csv --metrics $M python bench.py --ncu sample --steps 1 --warmup 3 --no-cpu --phase sample > gpurun_out/ncu_sample.log 2>&1
timeout 500 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_train.csv --metrics $M python bench.py --ncu train --steps 1 --warmup 3 --no-cpu --phase ppo > gpurun_out/ncu_train.log 2>&1
N="ncu --set full --clock-control none --import-source on"
( timeout 300 $N -k regex:attention_bwd -c 2 -f -o gpurun_out/r2d_attn_bwd python tests/prof_attention_shapes.py --once --bwd ) > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out/launches_*.csv gpurun_out/r2d_attn_bwd.ncu-rep; tail -n 2 gpurun_out/ncu_train.log
