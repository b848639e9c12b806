# final-ish single-GPU bench (with cpu baseline), reference arm, then profile of the default plan
timeout 900 python bench.py > gpurun_out/bench_12.json 2> gpurun_out/bench_12.err; echo "bench exit=$?"; cat gpurun_out/bench_12.json | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_12_ref.json 2> gpurun_out/bench_12_ref.err; echo "ref exit=$?"; cut -c1-300 gpurun_out/bench_12_ref.json
bash scripts/gpu_profile_wide.sh
