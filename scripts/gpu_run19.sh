mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_cli.py -m gpu -q -k "sex_chrom" > gpurun_out/pytest_gpu_19a.log 2>&1; echo "cli exit=$?"; grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_19a.log | tail -30
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sex_chrom or unknown_chrom" > gpurun_out/pytest_gpu_19b.log 2>&1; echo "parity exit=$?"; grep -E "^FAILED|passed|failed|^E  " gpurun_out/pytest_gpu_19b.log | tail -60
