set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "predict or pipeline or metrics or topn or near" > gpurun_out/pytest_exp.txt 2>&1; tail -3 gpurun_out/pytest_exp.txt
timeout 600 python bench.py --no-cpu-baseline --no-e2e --trace > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err; grep "scoring kernel\|merge + exact\|row kernels" gpurun_out/bench_exp.err; cut -c1-330 gpurun_out/bench_exp.json
