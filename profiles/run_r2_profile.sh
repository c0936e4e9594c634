set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.json 2> gpurun_out/launches_bench.err
wc -l gpurun_out/r2_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_predict_a32|k_fit_rows|k_gram_i8_tc' -s 4 -c 4 -f -o gpurun_out/r2_top python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
ls -la gpurun_out/
timeout 600 python bench.py --shape netflix --similarity conditional_probability --K 100 --trace > gpurun_out/bench_netflix_condprob.json 2> gpurun_out/bench_netflix_condprob.err; tail -20 gpurun_out/bench_netflix_condprob.err
timeout 900 python bench.py --shape large --K 100 --steps 2 --warmup 1 --no-cpu-baseline --trace > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; tail -20 gpurun_out/bench_large.err
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q --durations=10 > gpurun_out/pytest_full.txt 2>&1; tail -25 gpurun_out/pytest_full.txt
