#!/bin/bash
# round 2, call 19 (1 GPU): persistent full-grid stencil kernels with label prefetch (k_stencil_loop): parity, A/B per mode at 256^3 / 128^3 / 512^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/r2c19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c19_pytest.log; tail -4 gpurun_out/r2c19_pytest.log
for v in "GMG_STENCIL_LOOP=0" "GMG_STENCIL_LOOP=15" "GMG_STENCIL_LOOP=1" "GMG_STENCIL_LOOP=2" "GMG_STENCIL_LOOP=4" "GMG_STENCIL_LOOP=8"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c19_ab_$tag.json 2> gpurun_out/r2c19_ab_$tag.err; echo "$v rc=$?"
done
for v in "GMG_STENCIL_LOOP=0" "GMG_STENCIL_LOOP=15"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v GMG_TMA=0 timeout 300 python bench.py --workload vcycle --size 512 --steps 20 --warmup 5 > gpurun_out/r2c19_sweep_$tag.json 2> gpurun_out/r2c19_sweep_$tag.err; echo "sweep $v rc=$?"
  env $v timeout 300 python bench.py --quick --size 128 --steps 10 --warmup 3 > gpurun_out/r2c19_128_$tag.json 2> gpurun_out/r2c19_128_$tag.err; echo "128 $v rc=$?"
done
python scripts/show_bench.py gpurun_out/r2c19_ab_*.json gpurun_out/r2c19_sweep_*.json gpurun_out/r2c19_128_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|L0:|us x"
