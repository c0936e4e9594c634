set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in (4,2):
    d=json.load(open(f'gpurun_out/bench_n{n}.json')); print(n, {k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms')}, d['e2e']['value'])
PY
