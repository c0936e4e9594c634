set -x
./profiles/microbench/smem_atomics > gpurun_out/microbench_atomics.txt 2>&1
cat gpurun_out/microbench_atomics.txt
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -k "netflix" > gpurun_out/pytest_netflix.txt 2>&1; tail -15 gpurun_out/pytest_netflix.txt
timeout 2400 python -m pytest tests/test_gpu_fullsize.py -q -k "large" > gpurun_out/pytest_large.txt 2>&1; tail -15 gpurun_out/pytest_large.txt
