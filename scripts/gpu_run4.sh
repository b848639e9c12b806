timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wide or edge or emitted" > gpurun_out/pytest_gpu_4.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_4.log; tail -3 gpurun_out/pytest_gpu_4.log
for plan in "512,2,3" "256,4,3" "512,2,1"; do
  PM_WIDE_PLAN=$plan timeout 600 python bench.py --steps 4 --warmup 3 --sites-per-step 65536 --no-cpu-baseline > gpurun_out/bench_3_${plan//,/_}.json 2> gpurun_out/bench_3_${plan//,/_}.err; echo "plan $plan exit=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_3_${plan//,/_}.json"))
    print("$plan", round(d["value"]), "sites/s  kernel ms", d["roofline"]["kernel_ms_per_launch"], "frac", round(d["roofline"]["frac"],3), "evals/site", d["roofline"]["evaluations_per_site"], d["config"]["kernel_plan"][:60])
except Exception as e: print("fail", e)
PY
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_1.json 2> gpurun_out/bench_ref_1.err; echo "ref exit=$?"; cat gpurun_out/bench_ref_1.json | cut -c1-400
