set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 scripts/comm_bench.py 256 > gpurun_out/r15_comm_p2p.json 2> gpurun_out/r15_comm_p2p.err
GMG_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29912 scripts/comm_bench.py 256 > gpurun_out/r15_comm_nccl.json 2> gpurun_out/r15_comm_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29913 tests/mgpu_worker.py > gpurun_out/r15_worker_n2.log 2>&1
echo "rc=$?" >> gpurun_out/r15_worker_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29914 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r15_bench_n2.json 2> gpurun_out/r15_bench_n2.err
