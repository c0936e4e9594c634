set -x
ncu --set full --clock-control none --import-source on -k regex:k_fit_rows -s 1 -c 1 -o gpurun_out/prof_r2_fit python profiles/probe_config.py ml25m cosine 200 > gpurun_out/ncu_fit.log 2>&1
tail -3 gpurun_out/ncu_fit.log; ls -la gpurun_out/prof_r2_fit.ncu-rep
