mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
( timeout 300 $N -k regex:attention_bwd -c 2 -f -o gpurun_out/r2c_attn_bwd python tests/prof_attention_shapes.py --once --bwd ) > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out/r2c_attn_bwd.ncu-rep
