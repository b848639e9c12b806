run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --sites-per-step 65536 --no-cpu-baseline > gpurun_out/bench_14_$name.json 2> gpurun_out/bench_14_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_14_$name.json"))
    print("$name", round(d["value"]), "sites/s  kernel ms", round(d["roofline"]["kernel_ms_per_launch"],3), "frac", round(d["roofline"]["frac"],3), d["config"]["kernel_plan"][60:190])
except Exception as e: print("fail", e)
PY
}
run default PM_X=1
run u4_t256_lowregs PM_WIDE_PLAN=256,4,1 PM_WIDE_LOWREGS=1
run u4_t256 PM_WIDE_PLAN=256,4,1
run u2_t512_lowregs PM_WIDE_PLAN=512,2,1
