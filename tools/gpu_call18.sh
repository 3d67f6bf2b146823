set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q -k "tiling_invariance" > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
