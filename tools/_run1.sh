set -x
PT="python -m pytest -x -q -m gpu --timeout 180 --timeout-method thread"
timeout 300 $PT tests/test_probe_gpu.py tests/test_probe_multitask.py tests/test_sinks.py > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2i_tests.log
bash tools/run_ncu_towers.sh r2i
