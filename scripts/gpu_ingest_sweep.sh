# Host GLF ingest on the GPU box's cores: 3,000 uncompressed GLFs (1,000 trios), thread counts 1..all, three runs each
# (pm-tools ingest-bench; no GPU work).  (The round-2 sweep of the fill loop's row-block size and of non-temporal stores,
# profiles/r02_host_ingest_sweep.txt, used build-time switches that are gone.)
mkdir -p gpurun_out
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from polymutt_b200 import capi, glfio, synth
ped = synth.trios(1000); n = 20000
h, r = synth.generate_sites(ped, n, seed=20261018, device="cuda")
hdr = h.cpu().numpy().view(capi.SITE_HDR_DTYPE).reshape(-1); recs = r.cpu().numpy().view(capi.PERSON_SITE_DTYPE).reshape(n, ped.n_person)
t = time.time(); glfio.write_run_dir('/tmp/ing', ped, hdr, recs); print('wrote', round(time.time() - t, 1), 's')
PY
nproc
cd /tmp/ing
for thr in 1 4 8 16 0; do
  for i in 1 2 3; do
    echo -n "threads=$thr: "
    PM_TIMING=1 $GRAFT_REPO_ROOT/polymutt_b200/bin/pm-tools ingest-bench -p run.ped -d run.dat -g run.gif --batched $thr --batch 4096 2>&1 | grep -o 'decode [0-9.]* s\|fill [0-9.]* s\|sites_per_s": [0-9]*' | tr '\n' ' '; echo
  done
done
