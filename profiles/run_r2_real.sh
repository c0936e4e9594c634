set -x
timeout 900 python -m pytest tests/test_gpu_real.py -q -x > gpurun_out/pytest_real.txt 2>&1; tail -15 gpurun_out/pytest_real.txt
timeout 600 python profiles/probe_real.py ml25m cosine 200 > gpurun_out/probe_real.txt 2>&1; tail -20 gpurun_out/probe_real.txt
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_predict_a32' -s 2 -c 1 -f -o gpurun_out/r2_predict python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench2.json 2> gpurun_out/ncu_bench2.err
ls -la gpurun_out/*.ncu-rep
