# plain run first (must exit 0), then the ncu launch list (our kernels only) and one full capture of the dominant kernel
CMD="python bench.py --steps 2 --warmup 3 --sites-per-step 8192 --no-cpu-baseline --e2e-sites 512 --e2e-steps 1"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sites|k_compact|k_post" -c 60 --csv --log-file gpurun_out/launches_wide.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/prof_plain2.json 2> gpurun_out/prof_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:k_sites_wide -s 6 -c 1 -o gpurun_out/prof_wide $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
