set -x
timeout 400 python tools/tc_stream_probe.py 64 > gpurun_out/s14_stream_probe.txt 2>&1; grep "all \|no MMA  \|chunk=8\|no copies" gpurun_out/s14_stream_probe.txt
timeout 400 python tools/tc_stream_probe.py 256 > gpurun_out/s14_stream_probe256.txt 2>&1; grep "all \|no MMA  \|no copies" gpurun_out/s14_stream_probe256.txt
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/s14_pytest_tc.log 2>&1; tail -2 gpurun_out/s14_pytest_tc.log
