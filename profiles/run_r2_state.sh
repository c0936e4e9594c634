set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_gram_tc.py tests/test_gpu_ease.py -q -x > gpurun_out/pytest_a.txt 2>&1; tail -5 gpurun_out/pytest_a.txt
timeout 600 python bench.py --trace > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -60 gpurun_out/bench_n1.err
