set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python tools/lib_digest.py > gpurun_out/digest.log 2>&1
for v in evict; do ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_$v.so python tools/lib_digest.py >> gpurun_out/digest.log 2>&1; done
python tools/sweep.py set_massflux,omega,omega2,wvelocity,step2d_loop base evict > gpurun_out/sweep3.log 2>&1
du -sh gpurun_out
