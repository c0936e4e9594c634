set -x
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_gram_tc.py -q -x > gpurun_out/pytest_a.txt 2>&1; tail -5 gpurun_out/pytest_a.txt
SEL='test_predict_topn_matches_oracle or test_predict_lists_only'
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/sanitizer_racecheck2.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck2.log | tail -3
