set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s19_bench_c4_n2.json 2> gpurun_out/s19_bench_c4_n2.err; tail -c 300 gpurun_out/s19_bench_c4_n2.json
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/s19_bench_ref_n2.json 2> gpurun_out/s19_bench_ref_n2.err; tail -c 300 gpurun_out/s19_bench_ref_n2.json
timeout 300 python -m pytest tests/test_sharding_gloo.py -x -q > gpurun_out/s19_gloo.log 2>&1; tail -2 gpurun_out/s19_gloo.log
