run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --sites-per-step 65536 --no-cpu-baseline > gpurun_out/bench_15_$name.json 2> gpurun_out/bench_15_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_15_$name.json"))
    print("$name", round(d["value"]), "sites/s  kernel ms", round(d["roofline"]["kernel_ms_per_launch"],3), "frac", round(d["roofline"]["frac"],3), d["config"]["kernel_plan"][60:190])
except Exception as e: print("fail", e)
PY
}
run default PM_X=1
run u16_t64 PM_WIDE_PLAN=64,16,1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wide" 2>&1 | tail -2
PM_WIDE_PLAN=64,16,1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "trios1000" 2>&1 | tail -2
