#!/bin/bash
# usage: tools/ab2.sh <ngpus> <grid> "<tag>:<bench opts>" ...   (A/B of roms_b200_set_option switches under torchrun)
N=$1; G=$2; shift 2
for spec in "$@"; do
  tag=${spec%%:*}; opts=${spec#*:}
  if [ "$N" = 1 ]; then ROMS_B200_BENCH_OPTS="$opts" python bench.py --steps 20 --warmup 3 --grid $G --no-cpu --no-extras > gpurun_out/ab_${G}_n${N}_$tag.json 2> gpurun_out/ab_${G}_n${N}_$tag.err
  else ROMS_B200_BENCH_OPTS="$opts" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --grid $G --no-cpu --no-extras > gpurun_out/ab_${G}_n${N}_$tag.json 2> gpurun_out/ab_${G}_n${N}_$tag.err; fi
  python -c "
import json; d=json.load(open('gpurun_out/ab_${G}_n${N}_$tag.json')); print('$tag', round(d['ms_per_step'],4), 'loop', round(d['phase_ms']['step2d_loop'],4), 'sum', round(d['phase_ms_sum'],3), 'launches/step', d['gpu_launches']/20, d['state_digest'][:12])" || tail -5 gpurun_out/ab_${G}_n${N}_$tag.err
done
