mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -k "vcf or mixext" > gpurun_out/pytest_gpu_22.log 2>&1; echo "pytest exit=$?"; grep -E "^FAILED|passed|failed|^E   " gpurun_out/pytest_gpu_22.log | cut -c1-1500 | tail -30
