#!/bin/bash
# ncu evidence for the kernels of the full BENCHMARK cpp set (tools/phys_time.py runs 10 + 6 + 2 + 4 steps of it on BENCHMARK3)
set -x
python tools/phys_time.py > gpurun_out/r02_phys_time.json 2> gpurun_out/r02_phys_time.err || exit 1
cat gpurun_out/r02_phys_time.json
ncu --set full --import-source on --clock-control none -k regex:'k_(lmd_vmix|lmd_east|bulk_flux|bulk_stress|rho_eos|pre_step3d_t|t3dmix2_geo)' -s 20 -c 8 -f -o gpurun_out/r02_phys python tools/phys_time.py > gpurun_out/ncu_phys.log 2>&1
ls -la gpurun_out/r02_phys.ncu-rep
