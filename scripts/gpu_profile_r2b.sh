# Round-2 (second half) ncu evidence: the kernels added or changed after r02_ncu_*: the --in_vcf step (k_unpack_pl3,
# k_sites_wide one-warp plan, k_all_rows, k_post<CALLS>), k_unpack_wire on the default workload's host path, and the
# bi-allelic-only instance of the thread-per-site kernel (CEPH).  Every command first runs plainly.
mkdir -p gpurun_out
V="python bench.py --workload vcf200x5 --steps 2 --warmup 3 --sites-per-step 65536 --no-cpu-baseline --e2e-sites 16384 --e2e-steps 1"
$V > gpurun_out/profb_plain1.json 2> gpurun_out/profb_plain1.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|pm::k_|k_sites|k_post|k_unpack|k_compact|k_all_rows" -c 80 --csv --log-file gpurun_out/r02b_launches_vcf200x5.csv $V > gpurun_out/ncub_launch.log 2>&1
echo "vcf launch list exit=$?"
$V > gpurun_out/profb_plain2.json 2> gpurun_out/profb_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:k_post -s 6 -c 1 -o gpurun_out/r02b_post_vcf200x5 $V > gpurun_out/ncub_full_post.log 2>&1
echo "post capture exit=$?"
D="python bench.py --steps 2 --warmup 3 --sites-per-step 8192 --no-cpu-baseline --e2e-sites 2048 --e2e-steps 1"
$D > gpurun_out/profb_plain3.json 2> gpurun_out/profb_plain3.err && \
ncu --set full --clock-control none --import-source on -k regex:k_unpack_wire -c 1 -o gpurun_out/r02b_unpack_wire $D > gpurun_out/ncub_full_unpack.log 2>&1
echo "unpack capture exit=$?"
C="python bench.py --workload ceph20_ba --steps 2 --warmup 3 --sites-per-step 262144 --no-cpu-baseline --e2e-sites 512 --e2e-steps 1"
$C > gpurun_out/profb_plain4.json 2> gpurun_out/profb_plain4.err && \
ncu --set full --clock-control none --import-source on -k regex:k_sites_narrow -s 6 -c 1 -o gpurun_out/r02b_narrow_ceph20_ba $C > gpurun_out/ncub_full_narrow.log 2>&1
echo "narrow capture exit=$?"
ls -la gpurun_out/*.ncu-rep
