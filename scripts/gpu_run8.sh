run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --sites-per-step 65536 --no-cpu-baseline > gpurun_out/bench_8_$name.json 2> gpurun_out/bench_8_$name.err; echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_8_$name.json"))
    print("$name", round(d["value"]), "sites/s  kernel ms", round(d["roofline"]["kernel_ms_per_launch"],3), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), d["config"]["kernel_plan"][:150])
except Exception as e: print("fail", e)
PY
}
run g3_default PM_X=1
run g1_128x8_3blk PM_WIDE_PLAN=128,8,1 PM_WIDE_ONE_BUF=1
run g1_128x8_4blk PM_WIDE_PLAN=128,8,1 PM_WIDE_ONE_BUF=1 PM_WIDE_LOWREGS=1
run g1_128x8_2buf PM_WIDE_PLAN=128,8,1
run g1_256x4 PM_WIDE_PLAN=256,4,1 PM_WIDE_ONE_BUF=1
