timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wide or edge" > gpurun_out/pytest_gpu_7.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu_7.log; tail -3 gpurun_out/pytest_gpu_7.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --sites-per-step 65536 --no-cpu-baseline > gpurun_out/bench_7_$name.json 2> gpurun_out/bench_7_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_7_$name.json"))
    print("$name", round(d["value"]), "sites/s  kernel ms", round(d["roofline"]["kernel_ms_per_launch"],3), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), d["config"]["kernel_plan"][:80])
except Exception as e: print("fail", e)
PY
}
run default PM_X=1
run nokidtab PM_WIDE_NO_KIDTAB=1
run onebuf PM_WIDE_ONE_BUF=1
run t768 PM_WIDE_PLAN=256,4,3
tail -3 gpurun_out/bench_7_default.err
