mkdir -p gpurun_out
python scripts/gpu_shapes.py > gpurun_out/shapes_27.jsonl 2> gpurun_out/shapes_27.err; echo "shapes exit=$?"; cut -c1-130 gpurun_out/shapes_27.jsonl
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q 2>&1 | tail -3
