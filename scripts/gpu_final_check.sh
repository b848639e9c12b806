# Last check of the round on one GPU: the --in_vcf tests of the executable (its engine calls now go through the
# per-GPU sharding helper), smoke(), and the default bench line (cli_e2e now carries loop_value).
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_cli.py -m gpu -q -x -k "vcf" > gpurun_out/r3j_pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/r3j_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke exit=$?"
timeout 250 python bench.py > gpurun_out/r3j_bench_default.json 2> gpurun_out/r3j_bench_default.err; echo "bench exit=$?"; cut -c1-600 gpurun_out/r3j_bench_default.json
