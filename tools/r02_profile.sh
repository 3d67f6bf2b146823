#!/bin/bash
# Round-2 evidence run on one B200: tests and plain bench first (must exit 0), then the ncu passes of B200_PROFILING.md.
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || exit 1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
python bench.py --physics reduced --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_bench_n1_reduced.json 2> gpurun_out/r02_bench_n1_reduced.err
B="python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu --no-extras"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 180 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv $B > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_step2d -s 130 -c 2 -f -o gpurun_out/r02_step2d $B > gpurun_out/ncu_k1.log 2>&1
ncu --set full --clock-control none -k regex:'k_(pre_step3d|prsgrd|rhs3d|uv3dmix|step3d|omega|wvelocity|rho_eos|set_massflux|set_depth|lmd_vmix|bulk_flux|bulk_stress|t3dmix2_geo)' -s 17 -c 17 -f -o gpurun_out/r02_3d $B > gpurun_out/ncu_3d.log 2>&1
B1="python bench.py --grid b3tile8 --steps 2 --warmup 3 --spinup 2 --no-cpu --no-extras"
ncu --set full --clock-control none -k regex:k_step2d_loop -s 2 -c 1 -f -o gpurun_out/r02_step2d_loop $B1 > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out/*.ncu-rep
