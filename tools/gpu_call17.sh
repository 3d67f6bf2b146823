set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/rc.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 375 -c 150 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:k_step2d -s 300 -c 2 -o /tmp/full_s2d python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu > gpurun_out/ncu_full_s2d.log 2>&1
ncu -i /tmp/full_s2d.ncu-rep --page raw --csv > gpurun_out/full_s2d_raw.csv 2>/dev/null
du -sh gpurun_out
