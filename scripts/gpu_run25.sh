mkdir -p gpurun_out
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_25.json 2> gpurun_out/bench_25.err; echo "bench exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_25.json')); print(round(d['value']), 'kernel ms', round(d['roofline']['kernel_ms_per_launch'],3), 'frac', round(d['roofline']['frac'],4))"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wide or ceph or trios20" 2>&1 | tail -2
python scripts/gpu_shapes.py > gpurun_out/shapes_25.jsonl 2> gpurun_out/shapes_25.err; echo "shapes exit=$?"; cut -c1-160 gpurun_out/shapes_25.jsonl
