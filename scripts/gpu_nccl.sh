mkdir -p gpurun_out
N=${1:-4}
run() { tag=$1; shift; env "$@" NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV,TUNING timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/prof_allreduce.py $tag > gpurun_out/nccl_full_$tag.log 2>&1; grep -h "\[allreduce" gpurun_out/nccl_full_$tag.log > gpurun_out/nccl_$tag.txt; grep -h -i -E "nvls|NCCL_ALGO|NCCL_PROTO|nChannels|channels|Connected all" gpurun_out/nccl_full_$tag.log | sort | uniq -c | sort -rn | head -15 >> gpurun_out/nccl_$tag.txt; rm -f gpurun_out/nccl_full_$tag.log; }
run default X=1
run nvls NCCL_ALGO=NVLS
run ring NCCL_ALGO=Ring
run tree NCCL_ALGO=Tree
cat gpurun_out/nccl_*.txt
