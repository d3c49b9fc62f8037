set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sddmm_short -s 1 -c 1 -o gpurun_out/s8_prof_sddmm_short -f python tools/tc_one.py 1 16384 16384 64 2 0.1 sddmm > gpurun_out/s8_ncu.log 2>&1; tail -2 gpurun_out/s8_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_sddmm_short -s 1 -c 1 -o gpurun_out/s8_prof_sddmm_short_n256 -f python tools/tc_one.py 1 16384 16384 256 2 0.5 sddmm > gpurun_out/s8_ncu2.log 2>&1; tail -2 gpurun_out/s8_ncu2.log
