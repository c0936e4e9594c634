set -x
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_gram_tc.py -q -x > gpurun_out/pytest_a.txt 2>&1; tail -5 gpurun_out/pytest_a.txt
RPK_LIB=$PWD/recpack_b200/librpk_prof.so timeout 900 python profiles/probe_config.py ml25m cosine 200 > gpurun_out/probe_ml25m_prof.txt 2>&1; tail -32 gpurun_out/probe_ml25m_prof.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; python -c "
import json; d=json.load(open('gpurun_out/bench_a.json')); print({k:d[k] for k in ('value','ms_per_step','kernels_ms','phases_ms','ndcg10','recall20')}); print(d['e2e']['value'], d['e2e']['seconds'])"; tail -5 gpurun_out/bench_a.err
