set -x
# reference arm (short) and the default bench with cpu baseline
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms')}); print(d['cpu_baseline']); print(d['e2e']['value'])"; tail -5 gpurun_out/bench_n1.err
