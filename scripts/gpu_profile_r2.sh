# Round-2 ncu evidence: launch list of the default bench command, full captures of the block-per-site kernel (1,000 trios
# --denovo), the thread-per-site kernel (CEPH --denovo) and the posterior kernel.  Every command first runs plainly.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --sites-per-step 8192 --no-cpu-baseline --e2e-sites 512 --e2e-steps 1"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sites|k_compact|k_post|k_pack|k_quick" -c 60 --csv --log-file gpurun_out/r02_launches_trios1000_dn.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/prof_plain2.json 2> gpurun_out/prof_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:k_sites_wide -s 6 -c 1 -o gpurun_out/r02_wide_trios1000_dn $CMD > gpurun_out/ncu_full_wide.log 2>&1
echo "wide capture exit=$?"
CMD2="python bench.py --workload ceph20_dn --steps 2 --warmup 3 --sites-per-step 65536 --no-cpu-baseline --e2e-sites 512 --e2e-steps 1"
$CMD2 > gpurun_out/prof_plain3.json 2> gpurun_out/prof_plain3.err && \
ncu --set full --clock-control none --import-source on -k regex:k_sites_narrow -s 6 -c 1 -o gpurun_out/r02_narrow_ceph20_dn $CMD2 > gpurun_out/ncu_full_narrow.log 2>&1
echo "narrow capture exit=$?"
CMD3="python bench.py --workload ceph20_ba --steps 2 --warmup 3 --sites-per-step 262144 --no-cpu-baseline --e2e-sites 512 --e2e-steps 1"
$CMD3 > gpurun_out/prof_plain4.json 2> gpurun_out/prof_plain4.err && \
ncu --set full --clock-control none --import-source on -k regex:k_post -s 3 -c 1 -o gpurun_out/r02_post_ceph20_ba $CMD3 > gpurun_out/ncu_full_post.log 2>&1
echo "post capture exit=$?"
