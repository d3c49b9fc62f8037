set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s7_pytest_gpu.log 2>&1; tail -2 gpurun_out/s7_pytest_gpu.log
timeout 700 python benchmarks/run_configs.py --out gpurun_out/s7_configs.json > gpurun_out/s7_configs.txt 2> gpurun_out/s7_configs.err; tail -22 gpurun_out/s7_configs.txt | cut -c1-60,330-420
timeout 300 python benchmarks/fc_layer.py > gpurun_out/s7_fc_layers.txt 2> gpurun_out/s7_fc.err; tail -4 gpurun_out/s7_fc_layers.txt | cut -c1-300
