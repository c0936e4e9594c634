# compute-sanitizer over the GPU tests of the kernels added late in round 2 (real-valued fit, splitter, strips)
set -x
SEL='tests/test_gpu_real.py tests/test_gpu_split.py'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SEL -q -x -k "not larger_seeded" > gpurun_out/san_mem.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_mem.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest $SEL -q -x -k "unit or small or numpy_shuffles or golden" > gpurun_out/san_race.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/san_race.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "strips and not CASES2" > gpurun_out/san_strips.txt 2>&1; echo "memcheck strips rc=$?"; tail -4 gpurun_out/san_strips.txt
