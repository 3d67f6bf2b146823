#!/bin/bash
# bench.py under torchrun on N GPUs of one box, as the driver launches it.   tools/r02_scale.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print('N=$N', round(d['ms_per_step'],4), d['state_digest'][:16], 'loop', round(d['phase_ms']['step2d_loop'],4), 'weak', d.get('weak_scaling_row',{}).get('ms_per_step'), 'reduced', d.get('reduced_physics_row',{}).get('ms_per_step'), 'e2e', d['e2e']['value'])"
