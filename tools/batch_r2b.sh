set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2b_pytest_gpu.log
timeout 300 python tools/d2c_probe.py > gpurun_out/r2b_d2c_probe.txt 2>&1; tail -30 gpurun_out/r2b_d2c_probe.txt
timeout 200 python tools/dense2csr_run.py > gpurun_out/r2b_d2c_plain.txt 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:onepass -s 2 -c 1 -o gpurun_out/r2b_prof_d2c_onepass -f python tools/dense2csr_run.py > gpurun_out/r2b_ncu_d2c.log 2>&1; tail -3 gpurun_out/r2b_d2c_plain.txt
timeout 600 python benchmarks/run_configs.py --out gpurun_out/r2b_configs.json > gpurun_out/r2b_configs.txt 2> gpurun_out/r2b_configs.err; tail -22 gpurun_out/r2b_configs.txt
timeout 600 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; cat gpurun_out/r2b_bench_n1.json | head -c 3000
