mkdir -p gpurun_out
python scripts/gpu_shapes.py > gpurun_out/shapes_26.jsonl 2> gpurun_out/shapes_26.err; echo "shapes exit=$?"; cut -c1-130 gpurun_out/shapes_26.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ceph or ext or sex_chrom or vcf" 2>&1 | tail -2
