mkdir -p gpurun_out
for i in 1 2; do timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ab.json')); print(round(d['value']), 'kernel ms', round(d['roofline']['kernel_ms_per_launch'],3), 'frac', round(d['roofline']['frac'],4))"; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wide" 2>&1 | tail -2
