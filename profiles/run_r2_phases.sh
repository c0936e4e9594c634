set -x
RPK_LIB=$PWD/profiles/librpk_prof.so timeout 600 python profiles/prof_step.py > gpurun_out/phases_r2.txt 2>&1; tail -60 gpurun_out/phases_r2.txt
