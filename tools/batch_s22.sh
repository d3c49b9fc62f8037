set -x
for pf in 4 8 16; do
timeout 300 python benchmarks/run_configs.py --configs c3,c5 --no-dense --opt tc_prefetch=$pf --out gpurun_out/s22_configs_pf$pf.json > gpurun_out/s22_configs_pf$pf.txt 2> gpurun_out/s22_configs_pf$pf.err; echo "prefetch $pf"; grep "^| c3\|^| c5 16384^2 @[59]0" gpurun_out/s22_configs_pf$pf.txt | awk -F'|' '{print $2,"|",$4,"|",$5,"|",$13}'
done
