set -x
B200_TEST_OPTS=tc_pair=1 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/s16_pytest_pair.log 2>&1; tail -3 gpurun_out/s16_pytest_pair.log
timeout 400 python benchmarks/run_configs.py --configs c3,c5 --no-dense --opt tc_pair=1 --out gpurun_out/s16_configs_pair.json > gpurun_out/s16_configs_pair.txt 2> gpurun_out/s16_configs_pair.err; grep "^| c[35]" gpurun_out/s16_configs_pair.txt | awk -F'|' '{print $2,"|",$4,"|",$5,"|",$13,"|",$14}' | head -8
