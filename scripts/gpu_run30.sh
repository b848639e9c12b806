mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_30_n1.json 2> gpurun_out/bench_30_n1.err; echo "n1 exit=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_30_n2.json 2> gpurun_out/bench_30_n2.err; echo "n2 exit=$?"
python -c "
import json
for f in ('bench_30_n1.json','bench_30_n2.json'):
    d=json.load(open('gpurun_out/'+f)); print(f, d['n_gpus'], round(d['value']), d['e2e'])"
