set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 120 $TR tests/mgpu_check.py 512 64 30 8 > gpurun_out/mgpu2.log 2>&1; echo "mgpu rc=$?" > gpurun_out/rc.log
timeout 120 $TR tests/mgpu_check.py 96 40 30 5 > gpurun_out/mgpu2b.log 2>&1; echo "mgpu-b rc=$?" >> gpurun_out/rc.log
timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 200 $TR bench.py --gpus 2 --grid b3tile8x2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_t8.json 2> gpurun_out/bench_t8.err
