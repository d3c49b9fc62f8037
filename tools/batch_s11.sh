set -x
timeout 400 python tools/tc_stream_probe.py 64 > gpurun_out/s11_stream_probe.txt 2>&1; grep "all \|no MMA  \|chunk=8\|no copies" gpurun_out/s11_stream_probe.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s11_pytest_gpu.log 2>&1; tail -3 gpurun_out/s11_pytest_gpu.log
timeout 300 python tools/sddmm_short_probe.py > gpurun_out/s11_probe.txt 2>&1; tail -9 gpurun_out/s11_probe.txt | cut -c1-400
timeout 300 python tools/tc_breakdown.py > gpurun_out/s11_breakdown_c3.txt 2>&1; head -9 gpurun_out/s11_breakdown_c3.txt
