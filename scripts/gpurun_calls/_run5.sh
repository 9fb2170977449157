set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
PORT=29711
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r5_bench_n2.json 2> gpurun_out/r5_bench_n2.err
echo "rc=$?" >> gpurun_out/r5_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT+1)) tests/mgpu_worker.py > gpurun_out/r5_worker.log 2>&1
echo "rc=$?" >> gpurun_out/r5_worker.log
