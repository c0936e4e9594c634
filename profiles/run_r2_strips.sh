set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "strips or hybrid or shards" > gpurun_out/pytest_strips.txt 2>&1; tail -8 gpurun_out/pytest_strips.txt
timeout 900 python bench.py --shape large --K 100 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --trace > gpurun_out/bench_large2.json 2> gpurun_out/bench_large2.err; tail -40 gpurun_out/bench_large2.err; cat gpurun_out/bench_large2.json
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; cat gpurun_out/bench_n1b.json
