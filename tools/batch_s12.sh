set -x
timeout 200 python tools/tc_trace.py 16384 16384 64 0.1 > gpurun_out/s12_trace_c5.txt 2>&1; sed -n 300,345p gpurun_out/s12_trace_c5.txt
timeout 200 python tools/tc_trace.py 16384 16384 64 0.1 tc_terms=56 > gpurun_out/s12_trace_c5_nocopy_nomma.txt 2>&1; sed -n 300,330p gpurun_out/s12_trace_c5_nocopy_nomma.txt
