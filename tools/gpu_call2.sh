# full GPU test suite, bench A/B of the new knobs, kernel-variant sweep, fresh ncu evidence (single GPU).  Keeps gpurun_out/ small.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python bench.py > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "bench rc=$?" >> gpurun_out/rc.log
ROMS_B200_FUSE_TMIX=0 python bench.py --steps 10 --no-cpu > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
for mb in 48 96; do
  ROMS_B200_L2PERSIST=$mb python bench.py --steps 10 --no-cpu > gpurun_out/bench_l2_$mb.json 2> gpurun_out/bench_l2_$mb.err
done
python tools/lib_digest.py > gpurun_out/digest.log 2>&1
for v in ring4 ring6 pf4 s3u_d3m7; do ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_$v.so python tools/lib_digest.py >> gpurun_out/digest.log 2>&1; done
python tools/sweep.py pre_step3d,rhs3d,uv3dmix,step3d_uv,step3d_t pf4 pf8 s3u_d3m7 s3u_d4m6 ring4 ring6 ring8 > gpurun_out/sweep1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 152 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:k_step2d -s 300 -c 2 -o /tmp/full_s2d python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu > gpurun_out/ncu_full_s2d.log 2>&1
ncu -i /tmp/full_s2d.ncu-rep --page raw --csv > gpurun_out/full_s2d_raw.csv 2>/dev/null
ncu --set full --clock-control none -k 'regex:k_step3d|k_rhs3d|k_uv3dmix2|k_pre_step3d|k_omega|k_rho_eos|k_wvelocity|k_prsgrd|k_set_massflux|k_set_depth' -s 75 -c 16 -o /tmp/full_3d python bench.py --steps 2 --warmup 3 --spinup 2 --no-cpu > gpurun_out/ncu_full_3d.log 2>&1
ncu -i /tmp/full_3d.ncu-rep --page raw --csv > gpurun_out/full_3d_raw.csv 2>/dev/null
du -sh gpurun_out
