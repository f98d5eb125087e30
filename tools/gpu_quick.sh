set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/q_pytest.log
