mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_18.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_gpu_18.log
timeout 600 python bench.py > gpurun_out/bench_18_n1.json 2> gpurun_out/bench_18_n1.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_18_n1.json')); print(round(d['value']), d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
