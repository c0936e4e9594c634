set -x
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_all.txt 2>&1; tail -4 gpurun_out/pytest_all.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 600 python bench.py --shape netflix --similarity conditional_probability --K 100 --trace > gpurun_out/bench_netflix_condprob.json 2> gpurun_out/bench_netflix_condprob.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_netflix_condprob.json')); print({k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms')}); print(d['e2e']['value'], d['cpu_baseline']['value'])
PY
timeout 900 python bench.py --shape large --K 100 --steps 2 --warmup 1 --no-cpu-baseline --trace > gpurun_out/bench_large4.json 2> gpurun_out/bench_large4.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_large4.json')); print({k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms')}); print(d['e2e']['value'])
PY
