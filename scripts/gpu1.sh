mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_sd2_parity.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/t_sd2.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_sd2_parity.py 2>&1 | tail -40 > gpurun_out/t_all.log
DDPO_EXPERIMENTAL=1 DDPO_AESTHETIC_GPU=1 timeout 600 python -m pytest tests/test_gpu_zz_experimental.py -m gpu -q 2>&1 | tail -40 > gpurun_out/t_exp.log
timeout 900 python bench.py --shapes > gpurun_out/bench1.json 2> gpurun_out/bench1.err
DDPO_GN_REVERSE=1 timeout 300 python bench.py --phase sample --no-cpu > gpurun_out/bench_gnrev.json 2> gpurun_out/bench_gnrev.err
DDPO_GROUPED_TEMB=1 timeout 300 python bench.py --phase sample --no-cpu > gpurun_out/bench_gtemb.json 2> gpurun_out/bench_gtemb.err
timeout 300 python bench.py --phase sample --no-cpu > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
tail -3 gpurun_out/t_sd2.log gpurun_out/t_all.log gpurun_out/t_exp.log
