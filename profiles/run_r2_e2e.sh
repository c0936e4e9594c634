set -x
timeout 1500 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_split.py tests/test_gpu_ease.py tests/test_gpu_parity.py -q -x -k "not fit_" > gpurun_out/pytest_e2e.txt 2>&1; tail -4 gpurun_out/pytest_e2e.txt
timeout 900 python bench.py --trace > gpurun_out/bench_n1e.json 2> gpurun_out/bench_n1e.err; tail -3 gpurun_out/bench_n1e.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_n1e.json')); print({k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms')}); print(d['e2e']['value'], d['e2e']['seconds'], d['e2e']['step_seconds']); print(d['cpu_baseline']['value'])
PY
