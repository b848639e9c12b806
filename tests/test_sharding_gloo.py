"""The N>1 path on CPU: two gloo ranks each call the engine interface on their contiguous site range
(the CPU oracle stands in for the GPU engine here — this test is about the sharding logic), rank 0
concatenates the shards in rank order and must get exactly the single-process result; the timing
reduction takes the max over ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sites, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_lib import OracleEngine
    from polymutt_b200 import Params, capi, shard, synth
    ped = synth.concat(synth.trios(6), synth.families([4, 1, 1]))
    h, r = synth.generate_sites(ped, n_sites, seed=99, cfg=synth.SynthConfig(poly_boost=30.0))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    lo, hi = shard.site_range(n_sites, rank, world)
    eng = OracleEngine(ped, Params())
    st, res, per = eng.call_glf_sites(hdr[lo:hi], recs[lo:hi])
    eng.close()
    emitted = res[(st & 0xF) == 0].copy()
    emitted["site"] += lo                       # shard-local -> job-global site index
    merged = shard.gather_in_rank_order(emitted, dist)
    ms, total, rate = shard.reduce_timing(10.0 * (rank + 1), hi - lo, dist)
    if rank == 0:
        np.save(out_path, merged)
        assert ms == 10.0 * world and total == n_sites and abs(rate - n_sites / (ms * 1e-3)) < 1e-6
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(oracle_built, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import OracleEngine
    from polymutt_b200 import Params, capi, shard, synth
    n_sites = 1501
    # ranges tile the job, in order, for awkward sizes
    for world in (1, 2, 3, 8):
        r = [shard.site_range(n_sites, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == n_sites and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    out = str(tmp_path / "merged.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_sites, out), nprocs=2, join=True)
    merged = np.load(out)
    ped = synth.concat(synth.trios(6), synth.families([4, 1, 1]))
    h, r = synth.generate_sites(ped, n_sites, seed=99, cfg=synth.SynthConfig(poly_boost=30.0))
    hdr = h.numpy().view(capi.SITE_HDR_DTYPE).reshape(-1)
    recs = r.numpy().view(capi.PERSON_SITE_DTYPE).reshape(n_sites, ped.n_person)
    eng = OracleEngine(ped, Params())
    st, res, per = eng.call_glf_sites(hdr, recs)
    eng.close()
    want = res[(st & 0xF) == 0]
    assert len(want) > 20
    assert len(merged) == len(want)
    for name in want.dtype.names:   # field by field: pickling does not preserve struct padding bytes
        assert np.array_equal(merged[name], want[name], equal_nan=True), name
