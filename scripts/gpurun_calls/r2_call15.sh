#!/bin/bash
# round 2, call 15 (1 GPU): band sweep groups as ring-halo tiles (k_band_tile): parity tests, A/B at 256^3 / 512^3 / 128^3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiles or paths_agree" > gpurun_out/r2c15_pytest_a.log 2>&1; echo "pytest-a rc=$?" >> gpurun_out/r2c15_pytest_a.log; tail -15 gpurun_out/r2c15_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c15_pytest.log; tail -4 gpurun_out/r2c15_pytest.log
for v in "GMG_NONE=1" "GMG_BAND_TILES=0" "GMG_BAND_TILES_ZERO_ONLY=1"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v GMG_PRINT_STATS=1 timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/r2c15_ab_$tag.json 2> gpurun_out/r2c15_ab_$tag.err; echo "$v rc=$?"
  env $v timeout 300 python bench.py --workload vcycle --size 512 --steps 20 --warmup 5 > gpurun_out/r2c15_sweep_$tag.json 2> gpurun_out/r2c15_sweep_$tag.err; echo "sweep $v rc=$?"
  env $v timeout 300 python bench.py --quick --size 128 --steps 10 --warmup 3 > gpurun_out/r2c15_128_$tag.json 2> gpurun_out/r2c15_128_$tag.err; echo "128 $v rc=$?"
done
grep -h "band tiles" gpurun_out/r2c15_ab_GMG_NONE_1.* | sort | uniq | head
python scripts/show_bench.py gpurun_out/r2c15_ab_*.json gpurun_out/r2c15_sweep_*.json gpurun_out/r2c15_128_*.json 2>/dev/null | grep -E "==|value|vcycle_ms|roofline|L[0-9]:|us x"
