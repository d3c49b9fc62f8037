set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_spmm_kernel -s 2 -c 1 -o gpurun_out/s18_prof_tc_fwd_c5x64 -f python tools/tc_one.py 1 16384 16384 64 3 0.5 > gpurun_out/s18_ncu1.log 2>&1; tail -2 gpurun_out/s18_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_spmm_kernel -s 2 -c 1 -o gpurun_out/s18_prof_tc_fwd_c3 -f python tools/tc_one.py 1 3072 768 4096 3 > gpurun_out/s18_ncu2.log 2>&1; tail -2 gpurun_out/s18_ncu2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_spmm2_kernel -s 2 -c 1 -o gpurun_out/s18_prof_tc_pair_c5x256 -f python tools/tc_one.py 1 16384 16384 256 3 0.5 > gpurun_out/s18_ncu3.log 2>&1; tail -2 gpurun_out/s18_ncu3.log
