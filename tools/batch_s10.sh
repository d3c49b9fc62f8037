set -x
timeout 400 python tools/tc_stream_probe.py 64 > gpurun_out/s10_stream_probe.txt 2>&1; tail -45 gpurun_out/s10_stream_probe.txt
timeout 200 python tools/tc_trace.py 16384 16384 64 0.1 > gpurun_out/s10_trace_c5.txt 2>&1; sed -n 200,260p gpurun_out/s10_trace_c5.txt
