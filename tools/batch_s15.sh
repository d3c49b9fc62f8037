set -x
timeout 400 python benchmarks/run_configs.py --configs c5 --no-dense --opt tc_pair=1 --out gpurun_out/s15_configs_pair.json > gpurun_out/s15_configs_pair.txt 2> gpurun_out/s15_configs_pair.err; grep "^| c5" gpurun_out/s15_configs_pair.txt | cut -c1-60,330-420
