set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s9_bench_c4_n2.json 2> gpurun_out/s9_bench_c4_n2.err; tail -c 300 gpurun_out/s9_bench_c4_n2.json
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload c5:0.9:64 --no-cpu --no-configs > gpurun_out/s9_bench_c5_90_64_n2.json 2> gpurun_out/s9_bench_c5_90_64_n2.err; tail -c 300 gpurun_out/s9_bench_c5_90_64_n2.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workload c5:0.9:64 --no-cpu --no-configs > gpurun_out/s9_bench_c5_90_64_n1.json 2> gpurun_out/s9_bench_c5_90_64_n1.err; tail -c 300 gpurun_out/s9_bench_c5_90_64_n1.json
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload c5:0.99:256 --no-cpu --no-configs > gpurun_out/s9_bench_c5_99_256_n2.json 2> gpurun_out/s9_bench_c5_99_256_n2.err; tail -c 300 gpurun_out/s9_bench_c5_99_256_n2.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workload c5:0.99:256 --no-cpu --no-configs > gpurun_out/s9_bench_c5_99_256_n1.json 2> gpurun_out/s9_bench_c5_99_256_n1.err; tail -c 300 gpurun_out/s9_bench_c5_99_256_n1.json
