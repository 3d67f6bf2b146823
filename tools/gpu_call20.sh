set -x
mkdir -p gpurun_out
python tools/sweep.py pre_step3d,uv3dmix,step3d_t base s3tm4 s3tm5 prtpp2 prtm2 prupp2 uvmm2 > gpurun_out/sweep7.log 2>&1
