timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu_6.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_6.log; tail -3 gpurun_out/pytest_gpu_6.log
for plan in "256,4,3" "128,8,3" "512,2,1"; do
  PM_WIDE_PLAN=$plan timeout 600 python bench.py --steps 4 --warmup 3 --sites-per-step 65536 --no-cpu-baseline > gpurun_out/bench_6_${plan//,/_}.json 2> gpurun_out/bench_6_${plan//,/_}.err; echo "plan $plan exit=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_6_${plan//,/_}.json"))
    print("$plan", round(d["value"]), "sites/s  kernel ms", d["roofline"]["kernel_ms_per_launch"], "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), d["config"]["kernel_plan"][:70])
except Exception as e: print("fail", e)
PY
done
tail -3 gpurun_out/bench_6_256_4_3.err
