mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong > gpurun_out/bench_n2_strong.json ) 2> gpurun_out/bench_n2_strong.err
tail -c 300 gpurun_out/bench_n2_strong.err; head -c 200 gpurun_out/bench_n2_strong.json
