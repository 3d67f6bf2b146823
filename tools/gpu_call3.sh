# new defaults (L2 prefetch, step2d row blocks, fused t3dmix): full GPU test suite, bench, variant sweep
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.log
python bench.py --steps 10 --no-cpu > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "bench rc=$?" >> gpurun_out/rc.log
python tools/lib_digest.py > gpurun_out/digest.log 2>&1
for v in nopf tring3 pf2 glue; do ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_$v.so python tools/lib_digest.py >> gpurun_out/digest.log 2>&1; done
python tools/sweep.py omega,wvelocity,pre_step3d,rhs3d,uv3dmix,step3d_uv,step3d_t,step2d_loop base nopf pf2 glue tring3 tring2 > gpurun_out/sweep2.log 2>&1
du -sh gpurun_out
