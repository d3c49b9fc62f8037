set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2c_bench_n1.json').read().strip().splitlines()[-1])
b=json.loads(open('gpurun_out/r2c_bench_n8.json').read().strip().splitlines()[-1])
print('N=1', a['value'], a['ms_per_step']); print('N=8', b['value'], b['ms_per_step'], 'ratio', b['value']/a['value'])
print(b['roofline'].get('kernel_ms_per_rank')); print({k:v for k,v in b.items() if k in ('gathered','phases')})
PY
