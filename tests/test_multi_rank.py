"""N > 1 path on CPU: world_size-2 gloo processes, each handling its contiguous pattern shard
(the 1-GPU run over shard k of G is the stand-in for rank k — SURVEY.md §4), then the optional
collation. Sharded + collated results must equal the unsharded run bit for bit."""
import os
import sys

import numpy as np
import pytest

from conftest import rib, ob, FlatCheck, mixed_patterns, ROOT


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest as cf
    from rindex_b200 import _shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text = cf.rib.gen_text("dna_drift", 120_000, 1_500, 3, 4242)
    N, m = 501, 9
    patt = cf.mixed_patterns(text, N, m, 17, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    host = cf.rib.HostIndex.from_text(text)
    engine = cf.FlatCheck(host, K=4)           # CPU stand-in for the per-rank GPU engine
    a, b = _shard.shard_bounds(N, world, rank)
    lo, hi, off, occ, _ = engine.locate(patt[a * m: b * m], b - a, m)
    glo, ghi = _shard.collate_ranges(lo, hi, N)
    goff, gocc = _shard.collate_occurrences(off, occ)
    np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), lo=glo, hi=ghi, off=goff, occ=gocc)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_run(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    text = rib.gen_text("dna_drift", 120_000, 1_500, 3, 4242)
    N, m = 501, 9
    patt = mixed_patterns(text, N, m, 17, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    port_idx = ob.PortIndex(text, sa=rib.suffix_array(text))
    elo, ehi, eoff, eocc, _ = port_idx.locate(patt, N, m)
    for r in range(2):
        g = np.load(str(tmp_path / ("rank%d.npz" % r)))
        assert np.array_equal(g["lo"], elo) and np.array_equal(g["hi"], ehi)
        assert np.array_equal(g["off"], eoff) and np.array_equal(g["occ"], eocc)


def test_shard_bounds_partition():
    from rindex_b200 import _shard
    for N in (0, 1, 7, 100, 1001):
        for G in (1, 2, 3, 8):
            cuts = [_shard.shard_bounds(N, G, r) for r in range(G)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(G - 1))
