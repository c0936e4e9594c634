set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.json 2> gpurun_out/launches_bench.err
tail -2 gpurun_out/launches_bench.err; wc -l gpurun_out/r2_launches.csv
