for t in 7 15 55 39 23 0 8; do echo "terms=$t"; python tools/tc_epi_probe.py "only=@90% x64" tc_terms=$t 2>&1 | grep "c5"; done > gpurun_out/s3_epi_terms.txt 2>&1
cat gpurun_out/s3_epi_terms.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_spmm_kernel -s 1 -c 1 -o gpurun_out/s3_prof_tc_sddmm -f python tools/tc_one.py 1 16384 16384 64 1 0.1 sddmm > gpurun_out/s3_ncu_tc_sddmm.log 2>&1; tail -2 gpurun_out/s3_ncu_tc_sddmm.log
