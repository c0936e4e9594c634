set -x
timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_gram_tc.py tests/test_gpu_ease.py -q -x > gpurun_out/pytest_a.txt 2>&1; tail -8 gpurun_out/pytest_a.txt
