set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x > gpurun_out/pytest_bal.txt 2>&1; tail -4 gpurun_out/pytest_bal.txt
timeout 600 python profiles/probe_shards.py 8 > gpurun_out/probe_shards8b.txt 2>&1; tail -9 gpurun_out/probe_shards8b.txt
timeout 600 python bench.py --no-cpu-baseline --no-e2e --trace > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; tail -18 gpurun_out/bench_n1c.err; cut -c1-400 gpurun_out/bench_n1c.json
