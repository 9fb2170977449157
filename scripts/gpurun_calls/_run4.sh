set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_facade.py "tests/test_gpu_parity.py::test_diagonal_free_plain_cg_matches_oracle_operators" "tests/test_gpu_properties.py::test_config4_liquid_box_vcycle_contracts" -m gpu -q 2>&1 | tail -150 > gpurun_out/r4_tests.log
GMG_COARSE_FUSED=0 timeout 900 python -m pytest "tests/test_gpu_parity.py::test_diagonal_free_plain_cg_matches_oracle_operators" "tests/test_gpu_properties.py::test_config4_liquid_box_vcycle_contracts" -m gpu -q 2>&1 | tail -80 > gpurun_out/r4_tests_nofused.log
GMG_NO_GRAPHS=1 timeout 900 python -m pytest "tests/test_gpu_parity.py::test_diagonal_free_plain_cg_matches_oracle_operators" "tests/test_gpu_properties.py::test_config4_liquid_box_vcycle_contracts" -m gpu -q 2>&1 | tail -80 > gpurun_out/r4_tests_nographs.log
