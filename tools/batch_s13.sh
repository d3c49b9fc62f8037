set -x
timeout 700 python benchmarks/run_configs.py --out gpurun_out/s13_configs.json > gpurun_out/s13_configs.txt 2> gpurun_out/s13_configs.err; tail -20 gpurun_out/s13_configs.txt | cut -c1-60,330-420
timeout 300 python benchmarks/fc_layer.py > gpurun_out/s13_fc_layers.txt 2> gpurun_out/s13_fc.err; tail -4 gpurun_out/s13_fc_layers.txt | cut -c1-300
timeout 600 python bench.py > gpurun_out/s13_bench_n1.json 2> gpurun_out/s13_bench_n1.err; tail -c 600 gpurun_out/s13_bench_n1.json
