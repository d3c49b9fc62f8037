set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s23_smoke.log 2>&1; tail -2 gpurun_out/s23_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s23_pytest_gpu.log 2>&1; tail -2 gpurun_out/s23_pytest_gpu.log
timeout 300 python bench.py --no-cpu > gpurun_out/s23_bench.json 2> gpurun_out/s23_bench.err; tail -c 300 gpurun_out/s23_bench.json
