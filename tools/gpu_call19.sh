set -x
mkdir -p gpurun_out
python tools/sweep.py rhs3d,uv3dmix,step3d_uv,step3d_t base ts32 ts128 bx32 s3tpf12 s3tpf4m5 > gpurun_out/sweep6.log 2>&1
