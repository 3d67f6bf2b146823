TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tests/mgpu_check.py 512 64 30 8 > gpurun_out/mgpu2.log 2>&1; grep -E "exchange path|MGPU_CHECK|MISMATCH|timed out" gpurun_out/mgpu2.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err
python -c "
import json
d=json.load(open('gpurun_out/bench_g1.json')); print('N1', d['ms_per_step'], d['e2e']['value'], 1e3*2048*256*30/d['e2e']['value'])"
run() { name=$1; shift; env "$@" $TR bench.py --gpus 2 --grid b3tile8x2 --steps 20 --warmup 3 --no-cpu > gpurun_out/t8_$name.json 2> gpurun_out/t8_$name.err; python -c "
import json,sys
d=json.load(open('gpurun_out/t8_$name.json')); p=d['phase_ms']; print('$name', round(d['ms_per_step'],3), 'e2e_ms', round(1e3*512*256*30/d['e2e']['value'],3), 's2d', round(p['step2d_loop'],3), '3d', round(sum(v for k,v in p.items() if k!='step2d_loop'),3))"; }
run default A=1
run nooverlap ROMS_B200_NO_OVERLAP=1
