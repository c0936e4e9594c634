set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "fit" > gpurun_out/pytest_fitq.txt 2>&1; tail -4 gpurun_out/pytest_fitq.txt
timeout 600 python bench.py --no-cpu-baseline --no-e2e --trace > gpurun_out/bench_n1d.json 2> gpurun_out/bench_n1d.err; grep "row kernels\|sort + values" gpurun_out/bench_n1d.err; cut -c1-330 gpurun_out/bench_n1d.json
