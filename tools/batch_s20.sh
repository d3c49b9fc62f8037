set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s20_pytest_gpu.log 2>&1; tail -2 gpurun_out/s20_pytest_gpu.log
timeout 400 python benchmarks/run_configs.py --no-dense --out gpurun_out/s20_configs.json > gpurun_out/s20_configs.txt 2> gpurun_out/s20_configs.err; grep "^| c" gpurun_out/s20_configs.txt | awk -F'|' '{print $2,"|",$4,"|",$5}'
timeout 300 python benchmarks/fc_layer.py > gpurun_out/s20_fc_layers.txt 2> gpurun_out/s20_fc.err; tail -2 gpurun_out/s20_fc_layers.txt | cut -c1-330
