set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_fit_rows|k_gram_i8_tc|k_fit_merge|k_fit_sort_rows|k_user_split|k_fill_csc|k_csc_prefix|k_heavy' --csv --log-file gpurun_out/r2_large_launches.csv python profiles/probe_config.py large cosine 100 > gpurun_out/probe_large_auto.txt 2>&1; tail -5 gpurun_out/probe_large_auto.txt
timeout 900 python profiles/probe_config.py large cosine 100 4096 > gpurun_out/probe_large_4096.txt 2>&1; tail -5 gpurun_out/probe_large_4096.txt
