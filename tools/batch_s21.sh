set -x
timeout 400 python tools/host_pipeline_check.py 20 > gpurun_out/s21_host_pipeline.txt 2>&1; cat gpurun_out/s21_host_pipeline.txt | tail -12
