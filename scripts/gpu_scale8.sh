mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/bench_28_n8.json 2> gpurun_out/bench_28_n8.err; echo "n8 exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_28_n8.json')); print(d['n_gpus'], round(d['value']), d['e2e']['value'], d['roofline']['frac'], d['scaling'])"; tail -3 gpurun_out/bench_28_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_28_ref8.json 2> gpurun_out/bench_28_ref8.err; echo "ref8 exit=$?"; cut -c1-160 gpurun_out/bench_28_ref8.json
