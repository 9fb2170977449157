set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN GMG_TRACE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29901 tests/mgpu_worker.py > gpurun_out/r14_worker_n2.log 2>&1
echo "rc=$?" >> gpurun_out/r14_worker_n2.log
unset GMG_TRACE
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29902 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r14_bench_n2.json 2> gpurun_out/r14_bench_n2.err
echo "rc=$?" >> gpurun_out/r14_bench_n2.err
GMG_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29903 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r14_bench_n2_nccl.json 2> gpurun_out/r14_bench_n2_nccl.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29904 bench.py --gpus 2 --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r14_sweep512_n2.json 2> gpurun_out/r14_sweep512_n2.err
