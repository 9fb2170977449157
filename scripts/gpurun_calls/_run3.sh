set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r3_smi.log 2>&1
export NCCL_DEBUG=WARN
PORT=29611
# (1) two-rank sharding worker, quick mode (own timeout: a deadlocked collective must not hang the box)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT tests/mgpu_worker.py quick > gpurun_out/r3_worker_quick.log 2>&1
echo "worker quick rc=$?" >> gpurun_out/r3_worker_quick.log
# (2) facade binary directly (segfault hunt) then the whole GPU suite
python - > gpurun_out/r3_facade.log 2>&1 <<'PY'
import sys, subprocess
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_facade as t
t.write_case('/tmp/case.bin', 'sphere', 32)
p = subprocess.run([t.BIN, '/tmp/case.bin'], capture_output=True, text=True)
print(p.returncode); print(p.stdout[-4000:]); print(p.stderr[-4000:])
PY
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r3_tests.log
# (3) bench at N=2 and N=1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3_bench_n2.json 2> gpurun_out/r3_bench_n2.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err
ls -la gpurun_out
