set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s3f_pytest_gpu.log 2>&1; tail -2 gpurun_out/s3f_pytest_gpu.log
timeout 700 python benchmarks/run_configs.py --out gpurun_out/s3f_configs.json > gpurun_out/s3f_configs.txt 2> gpurun_out/s3f_configs.err; tail -24 gpurun_out/s3f_configs.txt
timeout 300 python benchmarks/fc_layer.py > gpurun_out/s3f_fc_layers.txt 2> gpurun_out/s3f_fc.err; tail -12 gpurun_out/s3f_fc_layers.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_spmm_kernel -s 1 -c 1 -o gpurun_out/s3f_prof_tc_sddmm -f python tools/tc_one.py 1 16384 16384 64 1 0.1 sddmm > gpurun_out/s3f_ncu_tc_sddmm.log 2>&1; tail -1 gpurun_out/s3f_ncu_tc_sddmm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_spmm_kernel -s 2 -c 1 -o gpurun_out/s3f_prof_tc_fwd_c3 -f python tools/tc_one.py 1 3072 768 4096 3 > gpurun_out/s3f_ncu_tc_fwd.log 2>&1; tail -1 gpurun_out/s3f_ncu_tc_fwd.log
python tools/tc_trace.py > gpurun_out/s3f_trace_c3.txt 2>&1; tail -3 gpurun_out/s3f_trace_c3.txt
