#!/bin/bash
# A/B of the barotropic loop: persistent kernel (32x8 tiles, 2 CTAs/SM | 32x16 tiles, 1 CTA/SM) against one launch per call.
# usage: tools/loopk_ab.sh <ngpus> <grid>
N=${1:-1}; G=${2:-benchmark1}
run() { tag=$1; shift
  if [ "$N" = 1 ]; then env "$@" python bench.py --steps 20 --warmup 3 --grid $G --no-cpu --no-extras > gpurun_out/ab_${G}_n${N}_$tag.json 2> gpurun_out/ab_${G}_n${N}_$tag.err
  else env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --grid $G --no-cpu --no-extras > gpurun_out/ab_${G}_n${N}_$tag.json 2> gpurun_out/ab_${G}_n${N}_$tag.err; fi
  python -c "
import json; d=json.load(open('gpurun_out/ab_${G}_n${N}_$tag.json')); print('$tag', round(d['ms_per_step'],4), 'loop', round(d['phase_ms']['step2d_loop'],4), 'launches/step', d['gpu_launches']/20, d['state_digest'][:12])" || tail -5 gpurun_out/ab_${G}_n${N}_$tag.err
}
run k2_8 X=1
run k2_16 ROMS_B200_LIB=roms_trunk_mgh_b200/lib/var/libroms_b200_lk16.so
run k1 ROMS_B200_BENCH_OPTS=step2d_loop_kernel=0
