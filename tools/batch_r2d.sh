set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2d_pytest_gpu.log
timeout 200 python tools/prof_one.py c5_n16_slab > gpurun_out/r2d_plain_slab2.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_slab2 -s 1 -c 1 -o gpurun_out/r2d_prof_c5_n16_slab2 -f python tools/prof_one.py c5_n16_slab > gpurun_out/r2d_ncu_slab2.log 2>&1; tail -1 gpurun_out/r2d_ncu_slab2.log
timeout 600 python benchmarks/run_configs.py --out gpurun_out/r2d_configs.json > gpurun_out/r2d_configs.txt 2> gpurun_out/r2d_configs.err; tail -22 gpurun_out/r2d_configs.txt
timeout 300 python benchmarks/fc_layer.py > gpurun_out/r2d_fc_layers.txt 2> gpurun_out/r2d_fc.err; tail -12 gpurun_out/r2d_fc_layers.txt
