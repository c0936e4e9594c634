set -x
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n$N.json')); print($N, {k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms','ndcg10','recall20')}); print(d['e2e']['value'], d['e2e']['seconds'])"; tail -3 gpurun_out/bench_r2_n$N.err
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 1 --shape large --K 100 --no-e2e > gpurun_out/bench_r2_large_n8.json 2> gpurun_out/bench_r2_large_n8.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_large_n8.json')); print('large', {k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms','ndcg10','recall20')}); print(d['config'])"; tail -3 gpurun_out/bench_r2_large_n8.err
