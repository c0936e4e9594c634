set -x
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_dropin.py -q -x > gpurun_out/pytest_split.txt 2>&1; tail -8 gpurun_out/pytest_split.txt
timeout 600 python profiles/probe_split.py > gpurun_out/probe_split.txt 2>&1; tail -8 gpurun_out/probe_split.txt
timeout 900 python bench.py --shape large --K 100 --steps 2 --warmup 1 --no-cpu-baseline --trace > gpurun_out/bench_large3.json 2> gpurun_out/bench_large3.err; tail -12 gpurun_out/bench_large3.err; cut -c1-1500 gpurun_out/bench_large3.json
