#!/bin/bash
# A/B of tuning variants (tools/variant.py) of the physics kernels: per-phase times of the full BENCHMARK set
for v in base "$@"; do
  if [ "$v" = base ]; then unset ROMS_B200_LIB; else export ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_$v.so; fi
  python tools/phys_time.py 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['phase_ms']; print('$v', round(d['ms_per_step'],3), 't3dmix', p['t3dmix'], 'lmd', p['lmd_vmix'], 'eos', p['rho_eos'], 'pre', p['pre_step3d'], 'rhs', p['rhs3d'])"
done
