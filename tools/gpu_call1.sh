# round-1 re-validation + A/B of the two new knobs + fresh ncu evidence (single GPU)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python bench.py > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "bench rc=$?" >> gpurun_out/rc.log
ROMS_B200_FUSE_TMIX=0 python bench.py --steps 10 --no-cpu > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
for mb in 48 96; do
  ROMS_B200_L2PERSIST=$mb python bench.py --steps 10 --no-cpu > gpurun_out/bench_l2_$mb.json 2> gpurun_out/bench_l2_$mb.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 152 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -s 456 -c 76 -o gpurun_out/full_step python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/full_step.ncu-rep --page raw --csv > gpurun_out/full_step_raw.csv 2>/dev/null
ls -la gpurun_out
python tools/sweep.py pre_step3d,rhs3d,uv3dmix,step3d_uv,step3d_t pf4 pf8 s3u_d3m7 s3u_d4m6 > gpurun_out/sweep1.log 2>&1
