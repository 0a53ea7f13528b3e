"""N > 1 host logic on CPU: two gloo ranks shard the reads, verify their block (the CPU port stands in for the
device here -- tests may use oracle/), and rank 0 gathers exactly the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from floxer_b200 import gpu, sharding, synthetic
    from floxer_b200.batch import VerifyConfig, alignment_records
    from oracle import cpu_baseline
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    refs = [synthetic.random_reference(60_000, 3)]
    batch = synthetic.make_batch(refs, 9, 600, 0.06, 11, gpu.pex_build, seed_errors=1, decoy_fraction=0.3)
    cfg = VerifyConfig(interval_optimization=True)
    mine, first = sharding.shard(batch, rank, world)
    al, cg, stats = cpu_baseline.verify_reads(refs, mine, cfg, threads=1)
    gathered = sharding.gather_records(alignment_records(al, cg), first, dist)
    ms, units = sharding.reduce_timing(10.0 * (rank + 1), float(len(mine)), dist)
    if rank == 0:
        wal, wcg, _ = cpu_baseline.verify_reads(refs, batch, cfg, threads=1)
        want = alignment_records(wal, wcg)
        ok = gathered == want and ms == 10.0 * world and units == float(len(batch)) and len(want) > 0
        with open(out_path, "w") as f:
            f.write("ok" if ok else f"mismatch {len(gathered)} {len(want)} {ms} {units}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_process(tmp_path):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from floxer_b200 import gpu, sharding, synthetic
    refs = [synthetic.random_reference(30_000, 5)]
    batch = synthetic.make_batch(refs, 13, 400, 0.05, 2, gpu.pex_build, seed_errors=1)
    for world in (1, 2, 3, 8, 20):
        b = sharding.shard_bounds(batch, world)
        assert b[0][0] == 0 and b[-1][1] == len(batch)
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(lo <= hi for lo, hi in b)
