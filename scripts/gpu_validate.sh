mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_23.log 2>&1; echo "pytest exit=$?"; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_23.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_23_n1.json 2> gpurun_out/bench_23_n1.err; echo "bench exit=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_23_n1.json')); print(round(d['value']), d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference > gpurun_out/bench_23_ref.json 2> gpurun_out/bench_23_ref.err; echo "ref exit=$?"; cut -c1-200 gpurun_out/bench_23_ref.json
