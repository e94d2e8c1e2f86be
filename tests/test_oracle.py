"""CPU tier: pin the oracle.

The reference ships no golden outputs, so the oracle is pinned three ways:
  1. brute-force substring search on the text (SDSL-independent; the `ri-locate -c` idea,
     reference ri-locate.cpp:156-190) and an explicit suffix array for the locate ORDER;
  2. the reference's own code run here (oracle/_ref: unmodified reference headers over the
     SDSL-API shim), plus fixtures under tests/golden/ generated from it by make_golden.py;
  3. the known answers of SURVEY.md §4 for the bundled world_leaders dataset (slow; runs only
     when /root/reference/datasets is present and RINDEX_SLOW=1).
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import rib, ob, mixed_patterns, repetitive_text, needs_ref, GOLDEN


def _sa_bruteforce(t):
    s = bytes(t) + b"\0"
    return sorted(range(len(s)), key=lambda i: s[i:])


def test_port_matches_suffix_array_definition():
    """count = SA interval of P in T+\\0; locate_all = SA[hi], SA[hi-1], ..., SA[lo] (r_index.hpp:328-355)."""
    rng = np.random.default_rng(1)
    for it in range(60):
        n = int(rng.integers(1, 80))
        t = repetitive_text(n, int(rng.integers(1, 20)), 1, it, sigma=int(rng.choice([1, 2, 4])))
        sa = _sa_bruteforce(t)
        s = bytes(t) + b"\0"
        port = ob.PortIndex(t)
        N, m = 25, int(rng.integers(1, 5))
        patt = mixed_patterns(t, N, m, it)
        lo, hi, off, occ, _ = port.locate(patt, N, m)
        for p in range(N):
            P = bytes(patt[p * m:(p + 1) * m])
            rows = [x for x in range(len(sa)) if s[sa[x]:sa[x] + m] == P]
            if 0 in P or 1 in P:
                continue  # 0x00 never occurs (and is the brute-force sentinel here); 0x01 can only match
                          # the terminator row of the BWT — covered by the edge-case tests
            if rows:
                assert (int(lo[p]), int(hi[p])) == (rows[0], rows[-1])
                assert occ[int(off[p]):int(off[p + 1])].tolist() == [sa[x] for x in reversed(rows)]
            else:
                assert (int(lo[p]), int(hi[p])) == (1, 0)  # the empty range literal, r_index.hpp:175,184


def test_port_vs_bruteforce_counts():
    t = repetitive_text(5000, 200, 2, 5, sigma=4)
    port = ob.PortIndex(t)
    N, m = 200, 5
    patt = mixed_patterns(t, N, m, 9)
    lo, hi, off, occ, _ = port.locate(patt, N, m)
    for p in range(N):
        P = patt[p * m:(p + 1) * m]
        assert int(off[p + 1] - off[p]) == ob.brute_count(t, P)
        assert np.array_equal(np.sort(occ[int(off[p]):int(off[p + 1])]), ob.brute_locate_sorted(t, P))


@needs_ref
def test_reference_code_equals_port():
    """The reference's own r_index<> (over the shim) and the restatement agree on ranges, offsets,
    every occurrence (order included), single rank / Phi probes, and the logical index content."""
    rng = np.random.default_rng(2)
    for it in range(40):
        n = int(rng.integers(1, 500))
        t = repetitive_text(n, int(rng.integers(1, 60)), int(rng.integers(0, 4)), 100 + it, sigma=int(rng.choice([1, 2, 4, 15])))
        ref = ob.RefIndex.from_text(t)
        port = ob.PortIndex(t)
        a, b = ref.extract(), port.extract()
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
        N, m = 40, int(rng.integers(1, 7))
        patt = mixed_patterns(t, N, m, it)
        r1, r2 = ref.locate(patt, N, m), port.locate(patt, N, m)
        for x, y in zip(r1[:4], r2[:4]):
            assert np.array_equal(x, y)
        for _ in range(30):
            i = int(rng.integers(0, n + 2))
            c = int(rng.choice(np.unique(t)))
            assert ref.rank(i, c) == port.rank(i, c)
        # Phi over every SA value except SA[0] (= n-1... the text position whose Phi is undefined)
        sa = rib.suffix_array(t)
        for x in rng.integers(1, n + 1, size=20):
            assert ref.phi(int(sa[x])) == port.phi(int(sa[x])) == int(sa[x - 1])


def test_sa_checker_and_product_sais():
    """The product's SA-IS output passes the oracle's independent linear-time verifier; corrupted
    arrays are rejected."""
    for seed in range(5):
        t = repetitive_text(20000, 500, 3, seed, sigma=4)
        sa = rib.suffix_array(t)
        assert ob.port_lib().rio_check_sa(t.ctypes.data, t.size, sa.ctypes.data) == 0
        bad = sa.copy(); bad[[5, 6]] = bad[[6, 5]]
        assert ob.port_lib().rio_check_sa(t.ctypes.data, t.size, bad.ctypes.data) != 0
    t = rib.gen_text("versioned_doc", 50000, 1000, 96, 3)
    sa = rib.suffix_array(t)
    assert ob.port_lib().rio_check_sa(t.ctypes.data, t.size, sa.ctypes.data) == 0
    a = ob.PortIndex(t, sa=sa).extract()   # oracle over the supplied (verified) SA
    b = ob.PortIndex(t).extract()          # oracle over its own prefix-doubling sort
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def _golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")) if os.path.isdir(GOLDEN) else []


@pytest.mark.parametrize("fname", _golden_files())
def test_golden_fixtures(fname):
    """Fixtures produced by the REFERENCE code (tests/golden/make_golden.py ran oracle/_ref here)."""
    g = np.load(os.path.join(GOLDEN, fname))
    text, patt, N, m = g["text"], g["patterns"], int(g["N"]), int(g["m"])
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    lo, hi, off, occ, _ = port.locate(patt, N, m)
    assert np.array_equal(lo, g["lo"]) and np.array_equal(hi, g["hi"])
    assert np.array_equal(off, g["occ_offsets"])
    assert hashlib.sha256(occ.tobytes()).hexdigest() == str(g["occ_sha256"])
    if "occ" in g.files:
        assert np.array_equal(occ, g["occ"])
    ex = port.extract()
    assert int(g["r"]) == ex["r"]
    assert hashlib.sha256(b"".join(np.ascontiguousarray(ex[k]).tobytes() for k in
                                   ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"))).hexdigest() == str(g["index_sha256"])


@pytest.mark.slow
@pytest.mark.skipif(not (os.path.exists("/root/reference/datasets/texts.7z") and os.environ.get("RINDEX_SLOW") == "1"),
                    reason="bundled dataset check is slow (inflates a 1.9 GB solid 7z stream); set RINDEX_SLOW=1")
def test_known_answers_world_leaders():
    """SURVEY.md §4 / BASELINE.md §2: occ_t = 29,781,174 for world_leaders_1000_8.patt."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import sevenz
    text = sevenz.extract_one("/root/reference/datasets/texts.7z", "world_leaders")
    pf = sevenz.extract_one("/root/reference/datasets/patterns.7z", "world_leaders_1000_8.patt")
    N, m, patt = rib.parse_pattern_file(pf)
    assert (N, m) == (1000, 8)
    t = np.frombuffer(text, dtype=np.uint8)
    port = ob.PortIndex(t, sa=rib.suffix_array(t))
    lo, hi, _ = port.count(patt, N, m)
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
    assert int(nocc.sum()) == 29_781_174
    assert nocc[:5].tolist() == [298060, 306619, 25029, 22721, 252]
    assert hashlib.sha256(nocc.astype("<u8").tobytes()).hexdigest()[:16] == "f5e5ac89a564dacc"


@needs_ref
def test_reference_navigation_equals_suffix_array_truth():
    """Pins the oracle side of the navigation row: the reference's own operator[], LF(i), FL(i),
    F_at(i) and get_bwt (r_index.hpp:162-164, 224-271, 375-377) equal what an explicit suffix array gives."""
    from test_gpu_navigate import _truth
    for text in (rib.gen_text("dna_drift", 30_000, 700, 3, 17), np.frombuffer(b"abracadabra\xff\xfe" * 9, dtype=np.uint8)):
        sa = rib.suffix_array(text)
        bwt, lf, fl, f_at = _truth(text, sa)
        ref = ob.RefIndex.from_text(text)
        pos = np.arange(sa.size, dtype=np.uint64)
        assert np.array_equal(ref.navigate(0, pos), bwt.astype(np.uint64))
        assert np.array_equal(ref.navigate(1, pos), lf)
        assert np.array_equal(ref.navigate(2, pos), fl)
        assert np.array_equal(ref.navigate(3, pos), f_at.astype(np.uint64))
        assert np.array_equal(ref.get_bwt(), bwt)


@needs_ref
def test_reference_index_assembled_from_logical_arrays_equals_reference_built():
    """ref_from_logical (the reference's structure constructors over a given BWT + samples: bench.py's reference arm uses
    it when no reference-built .ri travelled) gives the index the reference's constructor builds: same logical content,
    same count / locate_all answers."""
    for text in [rib.gen_text("dna_drift", 150_000, 2_000, 3, 8), rib.gen_text("versioned_doc", 60_000, 2_000, 96, 2),
                 np.frombuffer(b"abracadabra", dtype=np.uint8), np.frombuffer(b"a", dtype=np.uint8)]:
        built = ob.RefIndex.from_text(text)
        host = rib.HostIndex.from_text(text)
        made = ob.RefIndex.from_logical(host.arrays())
        e1, e2 = built.extract(), made.extract()
        assert all(np.array_equal(e1[k], e2[k]) for k in ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"))
        N, m = 200, min(6, text.size)
        patt = mixed_patterns(text, N, m, 3)
        a, b = built.locate(patt, N, m), made.locate(patt, N, m)
        assert all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4]))
