# full GPU validation + whole-executable throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_16.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_gpu_16.log
timeout 600 python bench.py > gpurun_out/bench_16_n1.json 2> gpurun_out/bench_16_n1.err; echo "bench exit=$?"; cut -c1-600 gpurun_out/bench_16_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_16_ref.json 2> gpurun_out/bench_16_ref.err; echo "ref exit=$?"; cut -c1-300 gpurun_out/bench_16_ref.json
timeout 900 python scripts/gpu_cli_e2e.py 12000 10000 > gpurun_out/cli_e2e_16.jsonl 2> gpurun_out/cli_e2e_16.err; echo "cli e2e exit=$?"; cat gpurun_out/cli_e2e_16.jsonl
nproc
