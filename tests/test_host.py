"""CPU tier: host-side builder, container I/O, generators, pattern-file parsing."""
import os

import numpy as np
import pytest

from conftest import rib, ob, repetitive_text, needs_ref, GOLDEN


def test_builder_equals_oracle_content():
    """This repo's builder produces exactly the logical arrays the reference's construction defines
    (r_index.hpp:553-634 sufsort, :72-82 F, :108-146 pred/pred_to_run/samples_last)."""
    rng = np.random.default_rng(3)
    for it in range(40):
        n = int(rng.integers(1, 3000))
        t = repetitive_text(n, int(rng.integers(1, 200)), int(rng.integers(0, 5)), it, sigma=int(rng.choice([1, 2, 4, 15])))
        a = rib.HostIndex.from_text(t).arrays()
        b = ob.PortIndex(t).extract()
        assert a["n"] == b["n"] == n + 1 and a["r"] == b["r"]
        for k in ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"):
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
        assert int(a["run_lens"].sum()) == n + 1
        assert int(a["pred_pos"][-1]) == n  # last text position always sampled, r_index.hpp:129


@needs_ref
def test_builder_equals_reference_built_index():
    t = rib.gen_text("dna_drift", 120_000, 1_500, 3, 11)
    a = rib.HostIndex.from_text(t).arrays()
    b = ob.RefIndex.from_text(t).extract()
    for k in ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_reserved_characters_rejected():
    for bad in (b"abc\x00def", b"abc\x01def"):
        with pytest.raises(ValueError, match="reserved characters 0x0, 0x1"):
            rib.HostIndex.from_text(bad)


def test_save_load_roundtrip(tmp_path):
    t = repetitive_text(5000, 100, 2, 1)
    h = rib.HostIndex.from_text(t)
    for flag in (True, False):
        p = str(tmp_path / ("x%d.ri" % flag))
        h.save(p, with_flag_byte=flag)
        g = rib.HostIndex.load(p, with_flag_byte=flag)
        a, b = h.arrays(), g.arrays()
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    with pytest.raises(IOError):
        rib.HostIndex.load(str(tmp_path / "missing.ri"))
    junk = tmp_path / "junk.ri"
    junk.write_bytes(b"\x00not an index at all")
    with pytest.raises(IOError):
        rib.HostIndex.load(str(junk))
    # the leading byte is the reference's `fast` flag (ri-build.cpp:133): file = 1 byte + container
    assert open(str(tmp_path / "x1.ri"), "rb").read(9)[1:] == b"RIB200v1"


def test_generators_deterministic_and_shaped():
    a = rib.gen_text("dna_drift", 100_000, 1_000, 3, 42)
    b = rib.gen_text("dna_drift", 100_000, 1_000, 3, 42)
    c = rib.gen_text("dna_drift", 100_000, 1_000, 3, 43)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert set(np.unique(a).tolist()) == set(b"ACGT")
    d = rib.gen_text("versioned_doc", 50_000, 1_000, 96, 1)
    assert d.min() >= 0x0A and d.max() <= 0x7E and np.unique(d).size > 40
    e = rib.gen_text("pangenome", 60_000, 2_000, 50, 1)
    assert set(np.unique(e).tolist()) <= set(b"ACGTN") and (e == ord("N")).sum() >= 10
    f = rib.gen_text("dna_indep", 80_000, 2_000, 1_000_000, 1)
    assert f.size == 80_000
    # repetitive: far fewer runs than symbols
    h = rib.HostIndex.from_text(a)
    assert h.r * 10 < h.n
    p = rib.gen_patterns(a, 50, 7, 5)
    s = bytes(a)
    assert all(bytes(p[i * 7:(i + 1) * 7]) in s for i in range(50))
    assert np.array_equal(p, rib.gen_patterns(a, 50, 7, 5))


def test_pattern_file_format(tmp_path):
    """Pizza&Chili shape read by the reference (utils.hpp:57-91, ri-count.cpp:86-110): header line,
    then N*m raw bytes which may contain newlines and bytes >= 0x80."""
    body = np.frombuffer(b"ab\ncd\xff\x80ef" * 3, dtype=np.uint8)
    p = str(tmp_path / "p.patt")
    rib.write_pattern_file(p, body, 9, 3)
    N, m, got = rib.parse_pattern_file(open(p, "rb").read())
    assert (N, m) == (9, 3) and np.array_equal(got, body)
    N, m, _ = rib.parse_pattern_file(b"# number=7 length=10 file=genome.fasta forbidden=\n\t\n" + b"x" * 70)
    assert (N, m) == (7, 10)
    with pytest.raises(ValueError):
        rib.parse_pattern_file(b"# nomber=7 length=10 file=x forbidden=\n")
    with pytest.raises(ValueError):
        rib.parse_pattern_file(b"# number=7 length=10\n")  # no space after the last field


def _same_index(a, b):
    A, B = a.arrays(), b.arrays()
    assert (A["n"], A["r"]) == (B["n"], B["r"])
    for k in ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"):
        assert np.array_equal(A[k], B[k]), k


def test_pfp_builder_equals_sais_builder_small():
    """Scalable construction: prefix-free parsing gives exactly the arrays of the suffix-array
    route — tiny windows and moduli force every phrase shape (one-phrase texts, phrases of length w+1, triggers
    inside the first window, the final phrase holding the terminator), texts from empty to a few hundred bytes."""
    rng = np.random.default_rng(11)
    for it in range(400):
        n = int(rng.integers(0, 500))
        sigma = int(rng.choice([1, 2, 4, 20, 250]))
        base = rng.integers(2, 2 + sigma, size=max(1, int(rng.integers(1, 80))), dtype=np.uint8)
        t = np.resize(base, n).copy()
        for _ in range(int(rng.integers(0, 8))):
            if n:
                t[int(rng.integers(0, n))] = int(rng.integers(2, 2 + sigma))
        a = rib.HostIndex.from_text(t)
        b = rib.HostIndex.from_text_pfp(t, w=int(rng.integers(1, 7)), p=int(rng.integers(1, 14)))
        _same_index(a, b)
        assert b.pfp_stats["uniform_rows"] + b.pfp_stats["merged_rows"] == n + 1


@pytest.mark.parametrize("threads", ["2", "5", "13"])
def test_pfp_builder_threaded_paths(threads, monkeypatch):
    """The parse (trigger scan, phrase hashes) and the sweep (ranges cut at group boundaries) run on several host
    threads for large inputs; forced here on small ones, including more threads than phrases or groups."""
    monkeypatch.setenv("RIB_PFP_THREADS", threads)
    rng = np.random.default_rng(int(threads))
    for it in range(150):
        n = int(rng.integers(0, 700))
        sigma = int(rng.choice([1, 2, 4, 20]))
        base = rng.integers(2, 2 + sigma, size=max(1, int(rng.integers(1, 90))), dtype=np.uint8)
        t = np.resize(base, n).copy()
        for _ in range(int(rng.integers(0, 8))):
            if n:
                t[int(rng.integers(0, n))] = int(rng.integers(2, 2 + sigma))
        _same_index(rib.HostIndex.from_text(t), rib.HostIndex.from_text_pfp(t, w=int(rng.integers(1, 7)), p=int(rng.integers(1, 14))))
    t = rib.gen_text("pangenome", 1_500_000, 100_000, 2_000, 9)
    _same_index(rib.HostIndex.from_text(t), rib.HostIndex.from_text_pfp(t))


def test_pfp_builder_equals_sais_builder_generators_and_golden():
    for kind, args in (("dna_drift", (2_000_000, 20_000, 3, 7)), ("versioned_doc", (1_500_000, 5_000, 96, 8)),
                       ("pangenome", (1_500_000, 100_000, 2_000, 9)), ("dna_indep", (1_500_000, 50_000, 10_000, 10))):
        t = rib.gen_text(kind, *args)
        b = rib.HostIndex.from_text_pfp(t)
        _same_index(rib.HostIndex.from_text(t), b)
        assert b.pfp_stats["dict_bytes"] < t.size // 2          # the point of the method on repetitive input
    for fname in sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")):
        t = np.load(os.path.join(GOLDEN, fname))["text"]
        _same_index(rib.HostIndex.from_text(t), rib.HostIndex.from_text_pfp(t, w=4, p=11))
    with pytest.raises(ValueError):
        rib.HostIndex.from_text_pfp(np.frombuffer(b"ab\x01cd", dtype=np.uint8))
    small = rib.HostIndex.from_text_auto(np.frombuffer(b"abracadabra", dtype=np.uint8))
    assert small.used_pfp is False                                # below 16 MB the SA-IS route is used


def test_truncated_or_corrupt_index_file_is_refused(tmp_path):
    """rib::load checks n, r and the stream length before it allocates: a truncated file or a header with an absurd
    run count gives RIH_ERR_FORMAT (the CLIs print their error and exit), never a bad_alloc / terminate."""
    text = rib.gen_text("dna_drift", 50_000, 1_000, 3, 9)
    good = str(tmp_path / "good.rib")
    rib.HostIndex.from_text(text).save(good)
    assert rib.HostIndex.load(good).n == text.size + 1
    blob = open(good, "rb").read()
    o = blob.index(b"RIB200v1")      # the `fast` flag byte precedes the container (ri-build.cpp:133)
    assert int.from_bytes(blob[o + 8: o + 16], "little") == text.size + 1
    cases = {"cut_arrays": blob[: len(blob) // 2], "cut_header": blob[: o + 40], "empty": b"",
             "huge_r": blob[: o + 16] + (2**62).to_bytes(8, "little") + blob[o + 24:],
             "r_gt_n": blob[: o + 8] + (5).to_bytes(8, "little") + (9).to_bytes(8, "little") + blob[o + 24:],
             "bad_magic": blob[:o] + b"XXXXXXXX" + blob[o + 8:]}
    for name, data in cases.items():
        p = str(tmp_path / (name + ".rib"))
        open(p, "wb").write(data)
        with pytest.raises(Exception):
            rib.HostIndex.load(p)
