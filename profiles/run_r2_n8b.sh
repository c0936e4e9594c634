set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --trace > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -22 gpurun_out/bench_n8.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_n8.json')); print({k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms')}); print(d['e2e']['value'], d['e2e']['seconds'])
PY
