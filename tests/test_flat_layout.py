"""CPU tier: the flatten-at-load step (r-index_b200/csrc/flat_layout.hpp) checked with the scalar
test double (tests/support/flat_check.cpp), which mirrors the kernels' index arithmetic."""
import numpy as np
import pytest

from conftest import rib, ob, FlatCheck, mixed_patterns, repetitive_text, GOLDEN
import os


@pytest.mark.parametrize("K", [4, 8, 16])
def test_flat_walk_equals_oracle(K):
    rng = np.random.default_rng(K)
    for it in range(60):
        n = int(rng.integers(1, 1500))
        t = repetitive_text(n, int(rng.integers(1, 120)), int(rng.integers(0, 4)), 50 * K + it, sigma=int(rng.choice([1, 2, 4, 15])))
        host = rib.HostIndex.from_text(t)
        port = ob.PortIndex(t)
        N, m = 40, int(rng.integers(1, 8))
        patt = mixed_patterns(t, N, m, it)
        elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
        fc = FlatCheck(host, K=K, lf_log2=int(rng.choice([0, 1, 3, 6])), phi_log2=int(rng.choice([0, 1, 3, 6])))
        assert fc.rc == 0
        lo, hi, off, occ, chains = fc.locate(patt, N, m)
        assert np.array_equal(lo, elo) and np.array_equal(hi, ehi)
        assert np.array_equal(off, eoff) and np.array_equal(occ, eocc)
        assert chains >= int((ehi >= elo).sum())


@pytest.mark.parametrize("jump", [1, 2, 4, 8])
@pytest.mark.parametrize("wide", [False, True, 7])   # 32-bit words; 64-bit with 40-bit packed records and Phi entries; plain 64-bit
def test_jump_tables_equal_repeated_phi(jump, wide):
    """Phi^j (j = 1..D, composition of piecewise translations) == j applications of Phi, for every SA
    value it can legally be applied to, through the scalar table AND the bucket-record lookup (32- and
    64-bit words); locate with D occurrences per lookup reproduces the oracle."""
    rng = np.random.default_rng(10 * jump + wide)
    for it in range(25):
        n = int(rng.integers(1, 2500))
        t = repetitive_text(n, int(rng.integers(1, 150)), int(rng.integers(0, 4)), 7000 + 31 * jump + it, sigma=int(rng.choice([1, 2, 4, 15])))
        host = rib.HostIndex.from_text(t)
        fc = FlatCheck(host, K=4, phi_log2=int(rng.choice([0, 1, 4])), jump=jump, force_wide=wide)
        assert fc.rc == 0 and fc.jump == jump
        assert fc.phi_packed == (wide is True and jump == 4)   # 64-bit index, D = 4: 32-byte packed Phi entries
        assert host.r <= fc.lib.fc_pieces(fc.h) <= jump * host.r
        sa = rib.suffix_array(t)
        for x in range(jump, n + 1):  # SA[x] with at least D predecessors in SA order
            assert fc.lib.fc_check_jump(fc.h, int(sa[x])) == 0
        port = ob.PortIndex(t)
        N, m = 30, int(rng.integers(1, 5))
        patt = mixed_patterns(t, N, m, it)
        elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
        lo, hi, off, occ, _ = fc.locate(patt, N, m)
        assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(occ, eocc)


@pytest.mark.parametrize("seg", [1, 16, 32, 64, 256])
@pytest.mark.parametrize("wide", [False, True])
def test_seed_table_and_two_pass_expansion(seg, wide):
    """Seed table Phi^SEG (built by doubling) == SEG applications of Phi for every SA value it can legally
    be applied to, through the scalar table AND the 8-word bucket-record lookup; the two-pass expansion
    (chain heads + one seed per SEG-slot output window, then one work item per window) reproduces the
    oracle, gives every window exactly one seed and keeps every vector store aligned."""
    rng = np.random.default_rng(100 * seg + wide)
    for it in range(20):
        n = int(rng.integers(1, 4000))
        t = repetitive_text(n, int(rng.integers(1, 200)), int(rng.integers(0, 4)), 9000 + 17 * seg + it, sigma=int(rng.choice([1, 2, 4, 15])))
        host = rib.HostIndex.from_text(t)
        jump = int(rng.choice([1, 2, 4, 8] if wide else [1, 2, 4, 6, 8]))
        fc = FlatCheck(host, K=4, phi_log2=int(rng.choice([0, 1, 4])), jump=jump, force_wide=wide, seed_jump=seg)
        assert fc.rc == 0 and fc.seed_jump == (seg if seg > 1 else 0)
        if seg > 1:
            assert host.r <= fc.lib.fc_seed_pieces(fc.h) <= min(seg * host.r + 1, n + 1)
            sa = rib.suffix_array(t)
            for x in range(seg, n + 1):
                assert fc.lib.fc_check_seed(fc.h, int(sa[x])) == 0
        port = ob.PortIndex(t)
        for m in (1, 2, int(rng.integers(1, 6))):  # short patterns: long ranges, chains spanning many windows
            N = 30
            patt = mixed_patterns(t, N, m, it)
            elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
            lo, hi, off, occ, chains = fc.locate(patt, N, m)
            assert chains < 2**63
            assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(occ, eocc)


def test_seed_table_auto_policy():
    t = rib.gen_text("dna_drift", 300_000, 3_000, 3, 5)
    host = rib.HostIndex.from_text(t)
    assert FlatCheck(host).seed_jump == 128      # small index: the largest window size fits the budget
    assert FlatCheck(host, seed_jump=1).seed_jump == 0
    assert FlatCheck(host, seed_jump=48).rc == -1


def test_jump_table_auto_policy():
    t = rib.gen_text("dna_drift", 300_000, 3_000, 3, 5)
    fc = FlatCheck(rib.HostIndex.from_text(t))
    assert fc.jump == 4 and fc.w32       # small index: the Phi^1..4 table is L2-friendly
    assert FlatCheck(rib.HostIndex.from_text(t), jump=6).jump == 6            # six per 32-byte entry: on request, 32-bit words only
    assert FlatCheck(rib.HostIndex.from_text(t), force_wide=True).jump == 4
    assert FlatCheck(rib.HostIndex.from_text(t), jump=3).rc == -1
    assert FlatCheck(rib.HostIndex.from_text(t), jump=6, force_wide=True).rc == -1


def test_flat_walk_medium_texts():
    for kind, args in (("dna_drift", (300_000, 3_000, 3, 5)), ("versioned_doc", (200_000, 2_000, 96, 6)),
                       ("pangenome", (200_000, 4_000, 100, 7))):
        t = rib.gen_text(kind, *args)
        host = rib.HostIndex.from_text(t)
        port = ob.PortIndex(t, sa=rib.suffix_array(t))
        N, m = 400, 9
        patt = mixed_patterns(t, N, m, 3)
        elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
        lo, hi, off, occ, _ = FlatCheck(host, K=16).locate(patt, N, m)
        assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(occ, eocc)


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")))
def test_flat_walk_golden(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    host = rib.HostIndex.from_text(g["text"])
    lo, hi, off, occ, _ = FlatCheck(host, K=8).locate(g["patterns"], int(g["N"]), int(g["m"]))
    assert np.array_equal(lo, g["lo"]) and np.array_equal(hi, g["hi"]) and np.array_equal(off, g["occ_offsets"])
    if "occ" in g.files:
        assert np.array_equal(occ, g["occ"])


def test_flatten_rejects_invalid_indexes():
    t = repetitive_text(3000, 100, 2, 1)
    host = rib.HostIndex.from_text(t)  # arrays() are borrowed views: keep the owner alive while copying
    good = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in host.arrays().items()}
    assert FlatCheck(good).rc == 0

    def broken(**kw):
        d = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in good.items()}
        for k, f in kw.items():
            f(d[k]) if callable(f) else d.__setitem__(k, f)
        return FlatCheck(d).rc

    def bump(i):
        def f(a):
            a[i] += 1
        return f
    assert broken(run_lens=bump(0)) == -5           # lengths no longer sum to n
    assert broken(pred_pos=bump(-1)) == -5          # last sample must be n-1
    assert broken(F=bump(70)) == -5                 # F disagrees with the runs
    assert broken(n=good["n"] + 1) == -5

    def swap(a):
        a[[0, 1]] = a[[1, 0]]
    assert broken(pred_pos=swap) == -5              # pred must be ascending
    assert FlatCheck(good, K=5).rc == -1            # unsupported group size
