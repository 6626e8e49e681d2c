mkdir -p gpurun_out
N=${1:-4}
run() { tag=$1; shift; ( env "$@" NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV,TUNING timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/prof_allreduce.py $tag 2>&1 | grep -E "allreduce|NVLS|Channel|Algo|algo|channels|NCCL_ALGO|NCCL_PROTO|Connected|error|Error" | grep -v "Channel [0-9]*/[0-9]* :" | head -40 ) > gpurun_out/nccl_$tag.log; }
run default X=1
run nvls NCCL_ALGO=NVLS
run ring NCCL_ALGO=Ring
run ring32 NCCL_ALGO=Ring NCCL_MIN_NCHANNELS=32
run tree NCCL_ALGO=Tree
grep -h "allreduce" gpurun_out/nccl_*.log
