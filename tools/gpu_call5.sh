set -x
mkdir -p gpurun_out
python tools/sweep.py step2d_loop base ty16 ty6 ty4 ty10 > gpurun_out/sweep4.log 2>&1
for v in ty16 ty6 ty4; do ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_$v.so python tools/lib_digest.py 64 37 8 4 >> gpurun_out/digest4.log 2>&1; done
python tools/lib_digest.py 64 37 8 4 >> gpurun_out/digest4.log 2>&1
