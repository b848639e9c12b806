mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_20.log 2>&1; echo "pytest exit=$?"; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_20.log | tail -30
timeout 600 python bench.py > gpurun_out/bench_20_n1.json 2> gpurun_out/bench_20_n1.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_20_n1.json')); print(round(d['value']), d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'])"
bash scripts/gpu_profile_wide.sh
ls -la gpurun_out/*.ncu-rep
