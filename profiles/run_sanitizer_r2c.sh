set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tars.py -q -x -k "not 3000" > gpurun_out/san_tars_mem.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/san_tars_mem.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tars.py -q -x -k "200-64 or topn or tars_cosine_exp or bigK" > gpurun_out/san_tars_race.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/san_tars_race.txt
