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


def _worker_balanced(rank, world, port, tmpdir):
    """The strong-scaling flow of bench.py / GpuFleet::locate on CPU: count on equal-count shards, all-gather of the
    per-pattern counts, contiguous shards re-cut at equal occurrence mass, locate, collation."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest as cf
    from rindex_b200 import _shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text = cf.rib.gen_text("dna_drift", 150_000, 1_500, 3, 99)
    N, m = 403, 4                                    # short patterns: counts spread over three orders of magnitude
    patt = cf.mixed_patterns(text, N, m, 23, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    engine = cf.FlatCheck(cf.rib.HostIndex.from_text(text), K=4)
    a, b = _shard.shard_bounds(N, world, rank)
    per = max(_shard.shard_bounds(N, world, r)[1] - _shard.shard_bounds(N, world, r)[0] for r in range(world))
    lo, hi = engine.count(patt[a * m: b * m], b - a, m)
    cnt = torch.zeros(per, dtype=torch.int64)
    cnt[: b - a] = torch.from_numpy(np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.int64))
    allc = [torch.zeros(per, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, cnt)
    nocc = np.concatenate([allc[r].numpy()[: _shard.shard_bounds(N, world, r)[1] - _shard.shard_bounds(N, world, r)[0]] for r in range(world)])
    cuts = _shard.balanced_cuts(nocc, world)
    c0, c1 = cuts[rank], cuts[rank + 1]
    lo, hi, off, occ, _ = engine.locate(patt[c0 * m: c1 * m], c1 - c0, m)
    goff, gocc = _shard.collate_occurrences(off, occ)
    mass = [int(nocc[cuts[r]: cuts[r + 1]].sum()) for r in range(world)]
    eq = [int(nocc[_shard.shard_bounds(N, world, r)[0]: _shard.shard_bounds(N, world, r)[1]].sum()) for r in range(world)]
    np.savez(os.path.join(tmpdir, "bal%d.npz" % rank), off=goff, occ=gocc, cuts=np.array(cuts), mass=np.array(mass), eq=np.array(eq))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_mass_balanced_shards_equal_single_run(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker_balanced, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    text = rib.gen_text("dna_drift", 150_000, 1_500, 3, 99)
    N, m = 403, 4
    patt = mixed_patterns(text, N, m, 23, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    elo, ehi, eoff, eocc, _ = ob.PortIndex(text, sa=rib.suffix_array(text)).locate(patt, N, m)
    for r in range(2):
        g = np.load(str(tmp_path / ("bal%d.npz" % r)))
        assert np.array_equal(g["off"], eoff) and np.array_equal(g["occ"], eocc)   # contiguous cuts: concatenation = the single run
        assert g["cuts"][0] == 0 and g["cuts"][-1] == N
        assert max(g["mass"]) <= max(g["eq"])                                       # never worse than equal-count shards


def _worker_replicated(rank, world, port, tmpdir):
    """Replicated planning (rig_plan_batch_dev + rig_expand_shard_dev; bench.py --plan replicate) on CPU: every rank
    searches the WHOLE batch, derives the same cuts from the offsets it computed, and expands only its shard. No
    collective on the data path (the barrier below only keeps the processes together)."""
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest as cf
    from rindex_b200 import _shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text = cf.rib.gen_text("dna_drift", 150_000, 1_500, 3, 99)
    N, m = 403, 4
    patt = cf.mixed_patterns(text, N, m, 23, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    engine = cf.FlatCheck(cf.rib.HostIndex.from_text(text), K=4)
    lo, hi = engine.count(patt, N, m)                       # the whole batch, on every rank
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0))
    off = np.concatenate([[0], np.cumsum(nocc)]).astype(np.uint64)
    cuts = _shard.balanced_cuts(np.diff(off.astype(np.int64)).astype(np.uint64), world)   # from the offsets, as the device does
    c0, c1 = cuts[rank], cuts[rank + 1]
    _, _, soff, socc, _ = engine.locate(patt[c0 * m: c1 * m], c1 - c0, m)
    assert int(soff[-1]) == int(off[c1] - off[c0])          # the shard's slots are the batch's, rebased
    np.savez(os.path.join(tmpdir, "rep%d.npz" % rank), occ=socc, cuts=np.array(cuts), base=np.array([int(off[c0])]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicated_planning_equals_single_run(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker_replicated, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    text = rib.gen_text("dna_drift", 150_000, 1_500, 3, 99)
    N, m = 403, 4
    patt = mixed_patterns(text, N, m, 23, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    _, _, eoff, eocc, _ = ob.PortIndex(text, sa=rib.suffix_array(text)).locate(patt, N, m)
    g = [np.load(str(tmp_path / ("rep%d.npz" % r))) for r in range(2)]
    assert np.array_equal(g[0]["cuts"], g[1]["cuts"])                          # same cuts without talking
    assert int(g[0]["base"][0]) == 0 and int(g[1]["base"][0]) == g[0]["occ"].size
    assert np.array_equal(np.concatenate([g[0]["occ"], g[1]["occ"]]), eocc)    # shard outputs concatenate to the single run


def test_balanced_cuts_properties():
    from rindex_b200 import _shard
    rng = np.random.default_rng(3)
    for N in (0, 1, 2, 17, 1000):
        for G in (1, 2, 3, 8):
            nocc = (rng.pareto(1.1, size=N) * 50).astype(np.uint64)
            cuts = _shard.balanced_cuts(nocc, G)
            assert len(cuts) == G + 1 and cuts[0] == 0 and cuts[-1] == N
            assert all(x <= y for x, y in zip(cuts, cuts[1:]))
    # a batch whose whole mass sits in one pattern cannot be split: every other shard is (almost) empty
    nocc = np.zeros(100, dtype=np.uint64); nocc[40] = 10**9
    cuts = _shard.balanced_cuts(nocc, 4)
    assert sum(1 for a, b in zip(cuts, cuts[1:]) if a <= 40 < b) == 1


def test_cuts_from_offsets_equal_cuts_from_counts():
    """rig_plan_batch_dev finds the cuts by binary search in the offset array, rig_balanced_cuts_dev / GpuFleet by a
    running sum over the counts: one rule, two formulations — equal on skewed, sparse, all-zero and tiny batches."""
    from rindex_b200 import _shard
    rng = np.random.default_rng(5)
    for trial in range(400):
        N = int(rng.choice([0, 1, 2, 3, 7, 64, 1000]))
        kind = trial % 4
        if kind == 0:
            nocc = (rng.pareto(1.1, size=N) * 50).astype(np.uint64)
        elif kind == 1:
            nocc = np.zeros(N, dtype=np.uint64)
            if N:
                nocc[rng.integers(0, N, size=min(N, 3))] = rng.integers(1, 10**9, size=min(N, 3)).astype(np.uint64)
        elif kind == 2:
            nocc = np.zeros(N, dtype=np.uint64)
        else:
            nocc = rng.integers(0, 5, size=N).astype(np.uint64)
        off = np.concatenate([[0], np.cumsum(nocc)]).astype(np.uint64)
        for G in (1, 2, 3, 8, 64):
            for cost in (64, 1, 0):
                want = _shard.balanced_cuts(nocc, G, cost)
                assert _shard.balanced_cuts_from_offsets(off, G, cost) == want, (trial, N, G, cost)


def test_shard_bounds_partition():
    from rindex_b200 import _shard
    for N in (0, 1, 7, 100, 1001):
        for G in (1, 2, 3, 8):
            cuts = [_shard.shard_bounds(N, G, r) for r in range(G)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(G - 1))
