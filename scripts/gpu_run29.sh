mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_29_n2.json 2> gpurun_out/bench_29_n2.err; echo "n2 exit=$?"; wc -l gpurun_out/bench_29_n2.json; head -c 200 gpurun_out/bench_29_n2.json; echo; grep -c NCCL gpurun_out/bench_29_n2.err
