mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_properties.py -m gpu -q -k "quick or ceph or trios20 or many_units or unsupported or mixext" > gpurun_out/pytest_gpu_21.log 2>&1; echo "pytest exit=$?"; grep -E "^FAILED|passed|failed|^E   " gpurun_out/pytest_gpu_21.log | cut -c1-900 | tail -30
