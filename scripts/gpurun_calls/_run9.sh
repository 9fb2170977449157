set -x
mkdir -p gpurun_out
free -g > gpurun_out/r9_mem.log; nproc >> gpurun_out/r9_mem.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lazy or gauss" 2>&1 | tail -15 > gpurun_out/r9_tests.log
timeout 300 python bench.py --workload vcycle --size 256 --steps 10 --warmup 3 > gpurun_out/r9_sweep256.json 2> gpurun_out/r9_sweep256.err
timeout 900 python bench.py --workload vcycle --size 512 --steps 10 --warmup 3 > gpurun_out/r9_sweep512.json 2> gpurun_out/r9_sweep512.err
echo rc=$? >> gpurun_out/r9_sweep512.err
