set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.json 2> gpurun_out/launches_bench.err
wc -l gpurun_out/r2_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_predict_a32|k_fit_rows|k_gram_i8_tc' -s 5 -c 4 -f -o gpurun_out/r2_final python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.err
ls -la gpurun_out/*.ncu-rep
