set -x
timeout 600 python profiles/probe_shards.py 2 > gpurun_out/probe_shards2.txt 2>&1; tail -3 gpurun_out/probe_shards2.txt
timeout 600 python profiles/probe_shards.py 8 > gpurun_out/probe_shards8.txt 2>&1; tail -9 gpurun_out/probe_shards8.txt
