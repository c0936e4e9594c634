set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --trace > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print({k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms')})"; grep -A40 "trace of one step" gpurun_out/bench_n2.err
