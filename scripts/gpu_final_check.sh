# Last check of the round on one GPU: the --in_vcf --gpus test, smoke(), the CEPH --denovo bench line at its new step size.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_cli.py -m gpu -q -x -k "vcf_input_on_all_gpus or small_batches" > gpurun_out/r3h_pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/r3h_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke exit=$?"
timeout 200 python bench.py --workload ceph20_dn > gpurun_out/r3h_bench_ceph20_dn.json 2> gpurun_out/r3h_bench_ceph20_dn.err; echo "bench exit=$?"; cut -c1-900 gpurun_out/r3h_bench_ceph20_dn.json
