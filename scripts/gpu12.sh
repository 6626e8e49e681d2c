mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
( timeout 300 $N -k regex:attention_bwd_dkdv -c 1 -f -o gpurun_out/r2b_attn_dkdv python tests/prof_attention_shapes.py --once --bwd ) > gpurun_out/ncu_dkdv.log 2>&1
( timeout 300 $N -k regex:attention_bwd_dq -c 1 -f -o gpurun_out/r2b_attn_dq python tests/prof_attention_shapes.py --once --bwd ) > gpurun_out/ncu_dq.log 2>&1
( timeout 300 $N -k regex:attention_fwd_kernel -c 1 -f -o gpurun_out/r2b_attn_fwd python tests/prof_attention_shapes.py --once ) > gpurun_out/ncu_fwd.log 2>&1
( timeout 300 $N -k regex:igemm2 -c 4 -f -o gpurun_out/r2b_igemm2 python tests/prof_igemm_shapes.py --once ) > gpurun_out/ncu_igemm2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 3 gpurun_out/ncu_dkdv.log
