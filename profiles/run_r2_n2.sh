set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --trace > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -25 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
