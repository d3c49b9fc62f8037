set -x
timeout 600 python tools/sddmm_short_probe.py > gpurun_out/s5_probe.txt 2>&1; tail -14 gpurun_out/s5_probe.txt
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/s5_pytest_tc.log 2>&1; tail -3 gpurun_out/s5_pytest_tc.log
