set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; tail -2 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 400 gpurun_out/final_bench_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-comparators --no-configs > gpurun_out/final_ncu_launches.log 2>&1; tail -1 gpurun_out/final_ncu_launches.log | cut -c1-200
