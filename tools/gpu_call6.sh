set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tests/mgpu_check.py 512 64 30 8 > gpurun_out/mgpu2.log 2>&1; grep -E "exchange path|MGPU_CHECK|MISMATCH|timed out" gpurun_out/mgpu2.log
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
$TR bench.py --gpus 2 --grid b3tile8x2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_t8.json 2> gpurun_out/bench_t8.err
ROMS_B200_NO_OVERLAP=1 $TR bench.py --gpus 2 --grid b3tile8x2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_t8_noov.json 2> gpurun_out/bench_t8_noov.err
python bench.py --grid b3tile8x2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_t8_single.json 2> gpurun_out/bench_t8_single.err
