mkdir -p gpurun_out
df -h /tmp | tail -1; free -g | head -2
timeout 1500 python scripts/gpu_cli_e2e.py 60000 30000 12000 > gpurun_out/cli_e2e_17.jsonl 2> gpurun_out/cli_e2e_17.err; echo "cli e2e exit=$?"; cat gpurun_out/cli_e2e_17.jsonl; tail -5 gpurun_out/cli_e2e_17.err
for plan in 256,4,1 512,2,1; do PM_WIDE_PLAN=$plan timeout 300 python bench.py --steps 4 --warmup 3 --sites-per-step 65536 --no-cpu-baseline > gpurun_out/bench_17_$plan.json 2> gpurun_out/bench_17_$plan.err; echo "$plan exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_17_$plan.json')); print('$plan', round(d['value']), d['roofline']['frac'])"; done
