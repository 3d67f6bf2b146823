set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python tools/sweep.py step2d_loop base nos2pf > gpurun_out/sweep5.log 2>&1
