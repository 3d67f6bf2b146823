set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus 2 --grid b3tile4x2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_t4.json 2> gpurun_out/bench_t4.err
ROMS_B200_FUSED_XCHG=0 timeout 200 $TR bench.py --gpus 2 --grid b3tile4x2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_t4_nofused.json 2> gpurun_out/bench_t4_nofused.err
