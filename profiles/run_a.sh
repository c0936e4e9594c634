set -x
timeout 1800 python -m pytest tests/test_gpu_ease.py tests/test_gpu_dropin.py -q > gpurun_out/pytest_a.txt 2>&1; tail -30 gpurun_out/pytest_a.txt
