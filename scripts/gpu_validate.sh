# full GPU validation: tests, smoke, bench (default workload + the others), reference arm
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_n1.json'))
print(round(d['value']), round(d['e2e']['value']), d['roofline']['bound'], round(d['roofline']['frac'], 3), d['gpu_launches'], d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('parity_checked_sites'), d.get('cpu_baseline', {}).get('parity_rows_differing'), str(d.get('cpu_baseline', {}).get('sample'))[:160])
PY
for w in ceph20_ba ceph20_dn vcf200x5 mixed100; do
  timeout 600 python bench.py --workload $w --steps 4 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value']), round(d['e2e']['value']), d['roofline']['bound'], round(d['roofline']['frac'],3), round(d['roofline']['fp64']['frac'],3), round(d['roofline']['hbm']['frac'],3))" 2>&1 | tail -1
done
