# usage (from the repo root): bash scripts/gpurun_calls/_run_mgpu.sh N tag [sweep]
N=$1; TAG=$2; SWEEP=$3
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
PORT=$((29800 + N))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT tests/mgpu_worker.py quick > gpurun_out/${TAG}_worker_n$N.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_worker_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+20)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "rc=$?" >> gpurun_out/${TAG}_bench_n$N.err
if [ -n "$SWEEP" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+40)) bench.py --gpus $N --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/${TAG}_sweep512_n$N.json 2> gpurun_out/${TAG}_sweep512_n$N.err
echo "rc=$?" >> gpurun_out/${TAG}_sweep512_n$N.err
fi
