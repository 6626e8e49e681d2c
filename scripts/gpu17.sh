mkdir -p gpurun_out
export DDPO_ALLOW_STUB_REWARDS=1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json ) 2> gpurun_out/bench_n2.err
tail -c 400 gpurun_out/bench_n2.err; head -c 300 gpurun_out/bench_n2.json
