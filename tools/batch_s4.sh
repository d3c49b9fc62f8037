set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s4_pytest_gpu.log 2>&1; tail -2 gpurun_out/s4_pytest_gpu.log
timeout 700 python benchmarks/run_configs.py --out gpurun_out/s4_configs.json > gpurun_out/s4_configs.txt 2> gpurun_out/s4_configs.err; tail -22 gpurun_out/s4_configs.txt | cut -c1-400
timeout 600 python bench.py > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err; tail -c 600 gpurun_out/s4_bench_n1.json
