set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python bench.py --no-cpu > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo "bench rc=$?" >> gpurun_out/rc.log
