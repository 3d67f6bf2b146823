set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR tests/mgpu_check.py 512 64 30 6 > gpurun_out/mgpu4.log 2>&1; echo "mgpu4 rc=$?" > gpurun_out/rc.log
timeout 200 $TR bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 rc=$?" >> gpurun_out/rc.log
