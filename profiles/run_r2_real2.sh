set -x
timeout 900 python -m pytest tests/test_gpu_real.py tests/test_gpu_tars.py -q -x > gpurun_out/pytest_real2.txt 2>&1; tail -3 gpurun_out/pytest_real2.txt
timeout 600 python profiles/probe_real.py ml25m cosine 200 > gpurun_out/probe_real2.txt 2>&1; tail -6 gpurun_out/probe_real2.txt
