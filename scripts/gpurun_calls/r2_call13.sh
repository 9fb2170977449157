#!/bin/bash
# round 2, call 13 (1 GPU): the whole GPU test-suite on the current library (mixed precision without a fused cycle, facade rebuilt), smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c13_pytest.log; tail -8 gpurun_out/r2c13_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c13_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c13_smoke.log
