# every BASELINE workload on 8 GPUs of one box (torchrun, one rank per GPU); bench lines land in gpurun_out/
mkdir -p gpurun_out
nvidia-smi -L | wc -l
port=29520
for w in trios1000_dn ceph20_ba ceph20_dn vcf200x5 mixed100; do
  port=$((port + 1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_${w}_n8.json 2> gpurun_out/r02_bench_final_${w}_n8.err; echo "$w n8 exit=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final_${w}_n8.json')); print('$w', d['n_gpus'], round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['scaling'])" 2>&1 | tail -1
done
