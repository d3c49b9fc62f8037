set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s17_pytest_gpu.log 2>&1; tail -3 gpurun_out/s17_pytest_gpu.log
timeout 700 python benchmarks/run_configs.py --configs c3,c5 --out gpurun_out/s17_configs.json > gpurun_out/s17_configs.txt 2> gpurun_out/s17_configs.err; grep "^| c[35]" gpurun_out/s17_configs.txt | awk -F'|' '{print $2,"|",$4,"|",$5,"|",$13,"|",$14}' | head -9
