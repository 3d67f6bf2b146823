# end-of-session sanity on one B200 (gpurun -- 'bash tools/final_check.sh'): GPU test suite, smoke(), one bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.log
python bench.py --no-cpu > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?" >> gpurun_out/rc.log
