mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_24_$name.json 2> gpurun_out/bench_24_$name.err; echo "$name exit=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_24_$name.json')); print('$name', round(d['value']), 'kernel ms', round(d['roofline']['kernel_ms_per_launch'],3), 'frac', round(d['roofline']['frac'],4))"
}
run constmut PM_X=1
run smemmut PM_NO_CONST_MUT=1
cp polymutt_b200/lib/libpolymutt_b200.so polymutt_b200/lib/new.so
cp polymutt_b200/lib/old_3feec1e.so polymutt_b200/lib/libpolymutt_b200.so
run old3feec1e PM_X=1
cp polymutt_b200/lib/new.so polymutt_b200/lib/libpolymutt_b200.so
run constmut_again PM_X=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wide_kernel_parity or quartets_denovo or mix_denovo" 2>&1 | tail -2
