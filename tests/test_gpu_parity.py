"""-m gpu: the CUDA path, called through the C ABI (include/rindex_gpu.h), against the oracle.

Bar: bit-exact (u64 ranges, offsets and every occurrence position, in locate_all order
SA[hi], SA[hi-1], ..., SA[lo] — reference internal/r_index.hpp:340-351)."""
import os

import numpy as np
import pytest

from conftest import rib, ob, mixed_patterns, repetitive_text, needs_ref

pytestmark = pytest.mark.gpu


def _check_all(gpu, oracle, patt, N, m, tag=""):
    elo, ehi, eoff, eocc, _ = oracle.locate(patt, N, m)
    lo, hi = gpu.count(patt, N, m)
    assert np.array_equal(lo, elo), tag + " count lo"
    assert np.array_equal(hi, ehi), tag + " count hi"
    lo2, hi2, off, occ = gpu.locate(patt, N, m)
    assert np.array_equal(lo2, elo) and np.array_equal(hi2, ehi), tag + " locate ranges"
    assert np.array_equal(off, eoff), tag + " offsets"
    assert np.array_equal(occ, eocc), tag + " occurrences"
    return eocc.size


@pytest.mark.parametrize("K", [4, 8, 16])
def test_wide_words_all_group_sizes(K, monkeypatch):
    """64-bit position words (as for n >= 2^32) with 40-bit packed block records, for every group size of the
    cooperative search kernel and the one-lane kernel."""
    monkeypatch.setenv("RIG_VARIANT", "8")
    rng = np.random.default_rng(900 + K)
    for it in range(15):
        n = int(rng.integers(1, 3000))
        text = repetitive_text(n, int(rng.integers(1, 200)), int(rng.integers(0, 4)), 7000 * K + it, sigma=int(rng.choice([1, 2, 4, 15, 200])))
        m = int(rng.integers(1, 9))
        patt = mixed_patterns(text, 64, m, it)
        gpu = rib.GpuIndex(rib.HostIndex.from_text(text), runs_per_block=K)
        assert gpu.info.words32 == 0
        _check_all(gpu, ob.PortIndex(text), patt, 64, m, "wide K=%d it=%d" % (K, it))
        gpu.close()


@pytest.mark.parametrize("K", [4, 8, 16])
def test_small_random_texts(K):
    rng = np.random.default_rng(100 + K)
    for it in range(40):
        n = int(rng.integers(1, 400))
        sigma = int(rng.choice([1, 2, 4, 15]))
        base = int(rng.integers(1, max(2, n // 3)))
        text = repetitive_text(n, base, int(rng.integers(0, 4)), 1000 * K + it, sigma=sigma)
        m = int(rng.integers(1, 7))
        N = 64
        patt = mixed_patterns(text, N, m, it)
        host = rib.HostIndex.from_text(text)
        port = ob.PortIndex(text)
        gpu = rib.GpuIndex(host, runs_per_block=K, lf_bucket_log2=int(rng.choice([0, 1, 5])),
                           phi_bucket_log2=int(rng.choice([0, 1, 5])))
        _check_all(gpu, port, patt, N, m, "K=%d it=%d" % (K, it))
        gpu.close()


@pytest.mark.parametrize("K", [4, 8, 16])
def test_repetitive_dna_medium(K):
    text = rib.gen_text("dna_drift", 400_000, 4_000, 3, 77)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    gpu = rib.GpuIndex(host, runs_per_block=K)
    for (N, m, seed) in [(1000, 8, 1), (777, 20, 2), (33, 1, 3), (500, 50, 4)]:
        patt = mixed_patterns(text, N, m, seed, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
        nocc = _check_all(gpu, port, patt, N, m, "dna N=%d m=%d" % (N, m))
        assert nocc > 0
    t = gpu.timing()
    assert t["launches"] >= 1 and t["occ_total"] > 0 and t["chains"] > 0


def test_large_alphabet_versioned_doc():
    text = rib.gen_text("versioned_doc", 300_000, 3_000, 96, 5)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    gpu = rib.GpuIndex(host)
    assert gpu.info.sigma > 60
    for (N, m, seed) in [(800, 30, 1), (400, 3, 2)]:
        patt = mixed_patterns(text, N, m, seed)
        _check_all(gpu, port, patt, N, m, "doc N=%d m=%d" % (N, m))


def test_edge_cases():
    text = np.frombuffer(b"abracadabra_abracadabra_\xff\xfe\xffabra" * 7, dtype=np.uint8)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text)
    gpu = rib.GpuIndex(host)
    # N = 0
    lo, hi = gpu.count(np.zeros(0, dtype=np.uint8), 0, 5)
    assert lo.size == 0 and hi.size == 0
    lo, hi, off, occ = gpu.locate(np.zeros(0, dtype=np.uint8), 0, 5)
    assert off.tolist() == [0] and occ.size == 0
    # m = 0: the full range, occ = n (terminator row included) — r_index.hpp:292-302 with an empty loop
    lo, hi, off, occ = gpu.locate(np.zeros(0, dtype=np.uint8), 3, 0)
    elo, ehi, eoff, eocc, _ = port.locate(np.zeros(0, dtype=np.uint8), 3, 0)
    assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(occ, eocc)
    assert int(hi[0] - lo[0] + 1) == text.size + 1
    # absent symbols, the terminator byte 0x01, 0x00 and 0xFF (F[256] sentinel)
    pats = [b"\x00", b"\x01", b"\xff", b"\xfe", b"z", b"a"]
    patt = np.frombuffer(b"".join(pats), dtype=np.uint8)
    _check_all(gpu, port, patt, len(pats), 1, "single bytes")
    lo, hi = gpu.count(patt, len(pats), 1)
    assert (lo[0], hi[0]) == (1, 0)          # 0x00 never occurs: the empty range is {1,0}
    assert hi[1] - lo[1] + 1 == 1            # 0x01 matches the terminator row only
    pats2 = [b"abra", b"\xff\xfe\xff\x61", b"a\x01ra", b"zzzz", b"dabr", b"_abr", b"ra_a", b"bra\xff"]
    patt2 = np.frombuffer(b"".join(pats2), dtype=np.uint8)
    _check_all(gpu, port, patt2, len(pats2), 4, "mixed")
    # pattern longer than the text
    patt3 = np.frombuffer(bytes(text) + b"a", dtype=np.uint8)
    _check_all(gpu, port, patt3, 1, patt3.size, "too long")
    # the whole text as a pattern occurs exactly once, at position 0
    lo, hi, off, occ = gpu.locate(text, 1, text.size)
    assert occ.tolist() == [0]


def test_single_run_and_tiny_texts():
    for t in [b"a", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"ab", b"ba", b"abababababababababab", b"\xff" * 40]:
        text = np.frombuffer(t, dtype=np.uint8)
        host = rib.HostIndex.from_text(text)
        port = ob.PortIndex(text)
        for K in (4, 16):
            gpu = rib.GpuIndex(host, runs_per_block=K)
            for m in (1, 2, 3):
                patt = mixed_patterns(text, 32, m, m)
                _check_all(gpu, port, patt, 32, m, repr(t[:8]))
            gpu.close()


def test_capacity_protocol_and_errors():
    text = rib.gen_text("dna_drift", 50_000, 1_000, 2, 9)
    host = rib.HostIndex.from_text(text)
    gpu = rib.GpuIndex(host)
    N, m = 100, 6
    patt = rib.gen_patterns(text, N, m, 4)
    import ctypes
    lo = np.zeros(N, dtype=np.uint64); hi = np.zeros(N, dtype=np.uint64); off = np.zeros(N + 1, dtype=np.uint64)
    tot = ctypes.c_uint64(0)
    small = np.zeros(3, dtype=np.uint64)
    rc = gpu.lib.rig_locate_batch(gpu.h, patt.ctypes.data, N, m, lo.ctypes.data, hi.ctypes.data, off.ctypes.data,
                                  small.ctypes.data, small.size, ctypes.byref(tot))
    assert rc == -4 and tot.value == off[-1] and tot.value > 3      # RIG_ERR_CAPACITY, ranges still filled
    assert np.array_equal(np.where(hi >= lo, hi - lo + 1, 0), np.diff(off))
    rc = gpu.lib.rig_count_batch(gpu.h, patt.ctypes.data, N, m, None, hi.ctypes.data)
    assert rc == -1                                                  # RIG_ERR_ARG
    assert gpu.lib.rig_strerror(-4) == b"occurrence buffer too small"
    # a corrupt index is rejected at create time
    bad = dict(host.arrays())
    bad["run_lens"] = bad["run_lens"].copy(); bad["run_lens"][0] += 1
    with pytest.raises(rib.RigError) as e:
        rib.GpuIndex(bad)
    assert e.value.code == -5
    with pytest.raises(rib.RigError) as e:
        rib.GpuIndex(host, device=99)
    assert e.value.code == -3


def test_device_buffer_api_and_digest():
    torch = pytest.importorskip("torch")
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 21)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    gpu = rib.GpuIndex(host)
    N, m = 2000, 10
    patt = rib.gen_patterns(text, N, m, 8)
    elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
    dev = torch.device("cuda:0")
    d_patt = torch.from_numpy(patt).to(dev)
    d_lo = torch.zeros(N, dtype=torch.int64, device=dev); d_hi = torch.zeros(N, dtype=torch.int64, device=dev)
    d_off = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    gpu.count_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_lo.cpu().numpy().view(np.uint64), elo)
    assert np.array_equal(d_hi.cpu().numpy().view(np.uint64), ehi)
    with pytest.raises(rib.RigError) as e:
        gpu.locate_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), None, 0, stream)
    assert e.value.code == -4 and e.value.needed == eocc.size
    d_occ = torch.zeros(eocc.size, dtype=torch.int64, device=dev)
    tot = gpu.locate_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), d_occ.data_ptr(),
                         d_occ.numel(), stream)
    torch.cuda.synchronize()
    assert tot == eocc.size
    assert np.array_equal(d_occ.cpu().numpy().view(np.uint64), eocc)
    assert np.array_equal(d_off.cpu().numpy().view(np.uint64), eoff)
    from rindex_b200._gpu import digest_host
    assert gpu.digest_dev(d_occ.data_ptr(), d_occ.numel(), stream) == digest_host(eocc)
    t = gpu.timing()
    assert t["occ_total"] == eocc.size and t["lf_steps"] > 0 and t["expand_ms"] > 0


@needs_ref
def test_against_reference_code_and_reference_built_index():
    """The reference's own r_index<> (oracle/_ref): same answers, and an index BUILT BY THE
    REFERENCE, extracted through its accessors (the INTEGRATION.md path), drives the GPU."""
    text = rib.gen_text("dna_drift", 250_000, 2_500, 3, 31)
    ref = ob.RefIndex.from_text(text)
    N, m = 1500, 12
    patt = mixed_patterns(text, N, m, 5, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    gpu = rib.GpuIndex(ref.extract())
    _check_all(gpu, ref, patt, N, m, "ref-built")
    gpu2 = rib.GpuIndex(rib.HostIndex.from_text(text))
    _check_all(gpu2, ref, patt, N, m, "own-built")


def test_full_size_properties_c2_like():
    """At a size where the oracle would take minutes: size-independent properties.
    n_occ == hi-lo+1; every reported position really holds the pattern (ri-locate -c idea,
    reference ri-locate.cpp:156-190); positions per pattern are distinct; count == locate ranges."""
    n = 20_000_000
    text = rib.gen_text("dna_drift", n, 50_000, 3, 0xB2000002)
    host = rib.HostIndex.from_text(text)
    gpu = rib.GpuIndex(host)
    N, m = 20_000, 20
    patt = rib.gen_patterns(text, N, m, 0xB2001002)
    lo, hi = gpu.count(patt, N, m)
    lo2, hi2, off, occ = gpu.locate(patt, N, m)
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0))
    assert np.array_equal(np.diff(off), nocc) and occ.size == int(nocc.sum())
    assert (nocc > 0).all()  # patterns were sampled from the text
    assert int(occ.max()) <= n - m
    P = patt.reshape(N, m)
    rng = np.random.default_rng(0)
    for p in rng.integers(0, N, size=300):
        o = occ[int(off[p]):int(off[p + 1])].astype(np.int64)
        assert np.unique(o).size == o.size
        win = text[o[:, None] + np.arange(m)[None, :]]
        assert (win == P[p][None, :]).all()
    # whole-output check for the first patterns: count by brute force
    for p in range(5):
        assert ob.brute_count(text, P[p]) == int(nocc[p])


@pytest.mark.parametrize("jump,variant", [(j, v) for j in (1, 2, 4, 6, 8) for v in ("3", "11", "1") if not (j == 6 and v == "11")])
def test_phi_jump_tables_and_wide_paths(jump, variant, monkeypatch):
    """Every expansion kernel variant gives the oracle's occurrences: D lanes per chain over the
    Phi^D jump table (D = 2, 4, 8), the one-lane-per-chain kernel (D = 1, coalesced and plain
    stores), and the 64-bit code paths used when n >= 2^32 (RIG_VARIANT bit 3)."""
    monkeypatch.setenv("RIG_VARIANT", variant)
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 99)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    gpu = rib.GpuIndex(host, phi_jump=jump)
    assert gpu.info.phi_jump == jump and gpu.info.words32 == (0 if variant == "11" else 1)
    for (N, m, seed) in [(1500, 9, 1), (300, 3, 2)]:
        patt = mixed_patterns(text, N, m, seed, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
        _check_all(gpu, port, patt, N, m, "jump=%d variant=%s" % (jump, variant))
    text2 = np.frombuffer(b"abracadabra" * 30, dtype=np.uint8)
    gpu2 = rib.GpuIndex(rib.HostIndex.from_text(text2), phi_jump=jump)
    patt = mixed_patterns(text2, 64, 2, 3)
    _check_all(gpu2, ob.PortIndex(text2), patt, 64, 2, "tiny jump=%d" % jump)


@pytest.mark.parametrize("seg", [1, 16, 64, 256])
@pytest.mark.parametrize("variant", ["0", "8", "32"])
def test_two_pass_expansion(seg, variant, monkeypatch):
    """Two-pass expansion (seed table Phi^SEG + one work item per SEG-slot output window) for every window
    size, 32- and 64-bit words, against the single-pass walk (SEG = 1, or RIG_VARIANT bit 5) and the oracle.
    Short patterns give ranges of 10^4..10^5 occurrences: chains that span hundreds of windows."""
    monkeypatch.setenv("RIG_VARIANT", variant)
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 123)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    for jump in ((4, 1) if variant == "8" else (6, 4, 1)):
        gpu = rib.GpuIndex(host, phi_jump=jump, seed_jump=seg)
        assert gpu.info.seed_jump == (seg if seg > 1 else 0)
        for (N, m, seed) in [(1200, 9, 1), (200, 2, 2), (40, 1, 3), (500, 30, 4)]:
            patt = mixed_patterns(text, N, m, seed, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
            _check_all(gpu, port, patt, N, m, "seg=%d jump=%d variant=%s" % (seg, jump, variant))
        t = gpu.timing()
        if seg > 1 and variant != "32":
            assert t["window_ms"] > 0 and t["seed_ms"] > 0
        gpu.close()


@pytest.mark.parametrize("variant", ["8192", "8200"])
def test_window_pass_alternatives(variant, monkeypatch):
    """The expansion as TWO kernels (RIG_VARIANT bit 13: seed pass, then window pass) instead of the default fused
    producer/consumer kernel — also the fallback when the fused kernel cannot be launched co-resident — in 32- and
    64-bit words (bit 3): the oracle's output."""
    monkeypatch.setenv("RIG_VARIANT", variant)
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 321)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    for jump, seg in ((4, 64), (1, 16), (8, 32), (2, 128)) + (((6, 128), (6, 16)) if variant == "8192" else ()):
        gpu = rib.GpuIndex(host, phi_jump=jump, seed_jump=seg)
        for (N, m, seed) in [(1200, 9, 1), (200, 2, 2), (40, 1, 3)]:
            patt = mixed_patterns(text, N, m, seed, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
            _check_all(gpu, port, patt, N, m, "window variant=%s jump=%d seg=%d" % (variant, jump, seg))
        gpu.close()


def test_two_pass_unaligned_output_falls_back_to_single_pass():
    """The window kernel's vector stores need a sector-aligned occurrence array; a caller-supplied device
    pointer that is only 8-byte aligned takes the single-pass walk and still gives the oracle's output."""
    torch = pytest.importorskip("torch")
    text = rib.gen_text("dna_drift", 200_000, 2_000, 3, 5)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    gpu = rib.GpuIndex(host)
    N, m = 800, 6
    patt = rib.gen_patterns(text, N, m, 8)
    elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
    dev = torch.device("cuda:0")
    d_patt = torch.from_numpy(patt).to(dev)
    d_lo = torch.zeros(N, dtype=torch.int64, device=dev); d_hi = torch.zeros(N, dtype=torch.int64, device=dev)
    d_off = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    buf = torch.zeros(eocc.size + 8, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for shift in (0, 1, 3):
        d_occ = buf[shift: shift + eocc.size]
        tot = gpu.locate_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(),
                             d_occ.data_ptr(), eocc.size, stream)
        torch.cuda.synchronize()
        assert tot == eocc.size and np.array_equal(d_occ.cpu().numpy().view(np.uint64), eocc)
        t = gpu.timing()
        assert (t["window_ms"] > 0) == (shift == 0)


@pytest.mark.parametrize("variant", ["128", "136", "8", "520", "648"])   # +512: 64-bit (not 40-bit packed) block records
def test_both_search_kernels_k4(variant, monkeypatch):
    """K = 4 block records are searched by one lane per pattern (default) or by the cooperative group kernel
    (RIG_VARIANT bit 7); both, in 32- and 64-bit words (bit 3), give the oracle's ranges, toeholds and occurrences,
    including early exits, absent symbols and patterns running off the text."""
    monkeypatch.setenv("RIG_VARIANT", variant)
    rng = np.random.default_rng(int(variant))
    for it in range(12):
        n = int(rng.integers(1, 6000))
        text = repetitive_text(n, int(rng.integers(1, 300)), int(rng.integers(0, 4)), 300 + it, sigma=int(rng.choice([1, 2, 4, 15])))
        gpu = rib.GpuIndex(rib.HostIndex.from_text(text), runs_per_block=4)
        port = ob.PortIndex(text)
        for m in (1, 3, int(rng.integers(4, 40))):
            patt = mixed_patterns(text, 100, m, it)
            _check_all(gpu, port, patt, 100, m, "variant=%s it=%d m=%d" % (variant, it, m))
        gpu.close()


def test_large_batch_capacity_protocol_and_caller_stream():
    """Batches of tens of thousands of patterns (hundreds of tiles in the search kernel's fused offset scan, several
    rounds of the persistent expansion kernels): same ranges, offsets and occurrences as the oracle; the capacity
    protocol (the expansion skips itself on the device when the buffer is too small) and the device-buffer entry point
    on a caller's stream behave the same; a second call on the same index reuses the grown item list."""
    torch = pytest.importorskip("torch")
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 41)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    gpu = rib.GpuIndex(host)
    for (N, m, seed) in [(16384, 12, 1), (20001, 7, 2), (40000, 16, 3), (129, 3, 4), (128, 2, 5)]:
        patt = mixed_patterns(text, N, m, seed, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
        elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
        for rep in range(2):
            lo, hi, off, occ = gpu.locate(patt, N, m)
            assert np.array_equal(lo, elo) and np.array_equal(hi, ehi) and np.array_equal(off, eoff) and np.array_equal(occ, eocc)
    import ctypes
    N, m = 20001, 7
    patt = mixed_patterns(text, N, m, 2, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    elo, ehi, eoff, eocc, _ = port.locate(patt, N, m)
    for cap in (3, eocc.size - 1):
        lo = np.zeros(N, dtype=np.uint64); hi = np.zeros(N, dtype=np.uint64); off = np.zeros(N + 1, dtype=np.uint64)
        small = np.zeros(cap, dtype=np.uint64)
        tot = ctypes.c_uint64(0)
        rc = gpu.lib.rig_locate_batch(gpu.h, patt.ctypes.data, N, m, lo.ctypes.data, hi.ctypes.data, off.ctypes.data,
                                      small.ctypes.data, small.size, ctypes.byref(tot))
        assert rc == -4 and tot.value == eocc.size and np.array_equal(off, eoff) and np.array_equal(lo, elo)
        assert not small.any()       # nothing was written into a buffer that is too small
    # device-buffer entry point on a caller's stream; the output buffer is poisoned first
    dev = torch.device("cuda:0")
    d_patt = torch.from_numpy(patt).to(dev)
    d_lo = torch.zeros(N, dtype=torch.int64, device=dev); d_hi = torch.zeros(N, dtype=torch.int64, device=dev)
    d_off = torch.zeros(N + 1, dtype=torch.int64, device=dev); d_occ = torch.full((eocc.size,), -1, dtype=torch.int64, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        tot = gpu.locate_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), d_occ.data_ptr(),
                             d_occ.numel(), s.cuda_stream)
    s.synchronize()
    assert tot == eocc.size and np.array_equal(d_occ.cpu().numpy().view(np.uint64), eocc)
    assert np.array_equal(d_off.cpu().numpy().view(np.uint64), eoff)


def test_item_list_regrows_when_chains_exceed_the_guess():
    """The item list of the two-pass expansion is sized before the totals are known (two chains per pattern + the
    capacity / SEG); short patterns on a text of many short runs cut every range into far more chains than that: the
    kernels skip themselves on the device and the host queues them again with a list that fits."""
    rng = np.random.default_rng(5)
    text = rng.integers(97, 101, size=120_000, dtype=np.uint8)   # iid: r ~ n, every range spans ~ its length in runs
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    gpu = rib.GpuIndex(host, seed_jump=16)
    assert gpu.info.seed_jump == 16
    for (N, m) in [(8, 1), (40, 2), (300, 3)]:
        patt = mixed_patterns(text, N, m, N, alphabet=np.frombuffer(b"abcd", dtype=np.uint8))
        _check_all(gpu, port, patt, N, m, "many chains N=%d m=%d" % (N, m))
        assert gpu.timing()["chains"] > 2 * N + 1024 or m == 3


def test_flat_index_file_round_trip(tmp_path):
    """rig_index_save_flat / rig_index_load_flat: the flattened index written to a file and loaded back (no flatten step)
    answers exactly like the index it was saved from, in 32- and 64-bit words; a file that belongs to another index,
    a truncated file and a file that is not a flat index are refused with RIG_ERR_INDEX / an error, never loaded."""
    import ctypes
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 77)
    host = rib.HostIndex.from_text(text)
    port = ob.PortIndex(text, sa=rib.suffix_array(text))
    N, m = 900, 7
    patt = mixed_patterns(text, N, m, 4, alphabet=np.frombuffer(b"ACGT", dtype=np.uint8))
    for variant in ("0", "8"):
        os.environ["RIG_VARIANT"] = variant
        try:
            gpu = rib.GpuIndex(host)
            path = str(tmp_path / ("idx%s.flat" % variant))
            gpu.save_flat(path)
            assert os.path.getsize(path) > gpu.info.device_bytes
            g2 = rib.GpuIndex(host, flat=path)
            assert g2.from_flat and g2.info.device_bytes == gpu.info.device_bytes and g2.info.seed_jump == gpu.info.seed_jump
            _check_all(g2, port, patt, N, m, "flat variant=%s" % variant)
            g3 = rib.GpuIndex(None, flat=path)     # no logical index at hand: the file is trusted
            _check_all(g3, port, patt, N, m, "flat (unchecked) variant=%s" % variant)
            # another index's flat file is refused, and the index is flattened the usual way instead
            other = rib.HostIndex.from_text(rib.gen_text("dna_drift", 200_000, 2_000, 3, 78))
            g4 = rib.GpuIndex(other, flat=path)
            assert not g4.from_flat and g4.n == other.n
            for g in (gpu, g2, g3, g4):
                g.close()
        finally:
            os.environ.pop("RIG_VARIANT", None)
    h = ctypes.c_void_p()
    lib = rib.gpu_lib()
    junk = tmp_path / "junk.flat"
    junk.write_bytes(b"not a flat index" * 100)
    assert lib.rig_index_load_flat(str(junk).encode(), None, 0, ctypes.byref(h)) == -5 and not h.value
    data = open(path, "rb").read()
    cut = tmp_path / "cut.flat"
    cut.write_bytes(data[: len(data) // 2])
    assert lib.rig_index_load_flat(str(cut).encode(), None, 0, ctypes.byref(h)) != 0 and not h.value
    assert lib.rig_index_load_flat(str(tmp_path / "absent.flat").encode(), None, 0, ctypes.byref(h)) == -1


def test_device_balanced_cuts_equal_host_rule():
    """rig_counts_dev + rig_balanced_cuts_dev (the multi-GPU fan-out helpers) give the cut points of
    _shard.balanced_cuts — the integer rule the C++ CLIs and the gloo test use — on skewed, empty and tiny batches."""
    torch = pytest.importorskip("torch")
    from rindex_b200 import _shard
    text = rib.gen_text("dna_drift", 50_000, 1_000, 3, 3)
    gpu = rib.GpuIndex(rib.HostIndex.from_text(text))
    rng = np.random.default_rng(11)
    dev = torch.device("cuda:0")
    for N in (1, 2, 9, 1000, 1025, 70_000):
        for W in (1, 2, 3, 8):
            nocc = (rng.pareto(1.2, size=N) * 100).astype(np.int64)
            if N > 5:
                nocc[rng.integers(0, N, size=3)] = 0
            d = torch.from_numpy(nocc).to(dev)
            assert gpu.balanced_cuts_dev(d.data_ptr(), N, W) == _shard.balanced_cuts(nocc.astype(np.uint64), W), (N, W)
    assert gpu.balanced_cuts_dev(0, 0, 4) == [0, 0, 0, 0, 0]
    lo = torch.tensor([5, 1, 0, 7], dtype=torch.int64, device=dev); hi = torch.tensor([9, 0, 0, 6], dtype=torch.int64, device=dev)
    out = torch.zeros(4, dtype=torch.int64, device=dev)
    gpu.counts_dev(lo.data_ptr(), hi.data_ptr(), 4, out.data_ptr())
    torch.cuda.synchronize()
    assert out.tolist() == [5, 0, 1, 0]


def test_plan_and_shard_expansion_equal_the_whole_batch(monkeypatch):
    """rig_plan_batch_dev + rig_expand_shard_dev (a job sharded over several GPUs, planned on every device): the cuts
    are the integer rule's, the ranges and offsets are the whole batch's, and the shards' occurrences, concatenated,
    are the whole batch's occurrences in locate_all order — for cut points of the plan and for arbitrary [c0, c1),
    over 32- and 64-bit words, fused / two-kernel / single-pass expansion."""
    torch = pytest.importorskip("torch")
    from rindex_b200 import _shard
    dev = torch.device("cuda:0")
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 23)
    host = rib.HostIndex.from_text(text)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for kw in (dict(), dict(seed_jump=16), dict(seed_jump=1), dict(phi_jump=2), dict(phi_jump=6)):
        gpu = rib.GpuIndex(host, **kw)
        for (N, m, seed) in [(900, 7, 1), (37, 1, 2), (1, 3, 3), (6000, 4, 4)]:
            patt = mixed_patterns(text, N, m, seed, alphabet=acgt)
            elo, ehi, eoff, eocc = gpu.locate(patt, N, m)
            nocc = np.diff(eoff.astype(np.int64)).astype(np.uint64)
            d_patt = torch.from_numpy(patt).to(dev)
            d_lo = torch.empty(N + 1, dtype=torch.int64, device=dev)
            d_hi = torch.empty(N + 1, dtype=torch.int64, device=dev)
            d_off = torch.empty(N + 2, dtype=torch.int64, device=dev)
            for W in (1, 2, 3, 8):
                cuts, total = gpu.plan_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), W)
                assert cuts == _shard.balanced_cuts(nocc, W), (kw, N, m, W)
                assert total == eocc.size
                assert np.array_equal(d_lo[:N].cpu().numpy().astype(np.uint64), elo) and np.array_equal(d_hi[:N].cpu().numpy().astype(np.uint64), ehi)
                assert np.array_equal(d_off[: N + 1].cpu().numpy().astype(np.uint64), eoff)
                bounds = list(zip(cuts, cuts[1:]))
                if W == 3 and N > 4:
                    bounds = [(0, 1), (1, N // 2), (N // 2, N // 2), (N // 2, N)]   # not cut points of the plan
                got = []
                for (c0, c1) in bounds:
                    with pytest.raises(rib.RigError) if eoff[c1] > eoff[c0] else _null():
                        gpu.expand_shard_dev(N, c0, c1, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), None, 0)
                    need = int(eoff[c1] - eoff[c0])
                    d_occ = torch.full((need + 16,), -1, dtype=torch.int64, device=dev)
                    assert gpu.expand_shard_dev(N, c0, c1, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), d_occ.data_ptr(), need) == need
                    torch.cuda.synchronize()
                    assert (d_occ[need:] == -1).all()
                    got.append(d_occ[:need].cpu().numpy().astype(np.uint64))
                assert np.array_equal(np.concatenate(got) if got else np.zeros(0, np.uint64), eocc), (kw, N, m, W)
    gpu.count(patt, N, m)                    # any other batch call invalidates the plan
    with pytest.raises(rib.RigError) as e:
        gpu.expand_shard_dev(N, 0, N, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), None, 0)
    assert e.value.code == -1
    monkeypatch.setenv("RIG_VARIANT", "8")   # 64-bit position words, as for n >= 2^32
    g64 = rib.GpuIndex(host)
    assert g64.info.words32 == 0
    if True:
        N, m = 500, 5
        patt = mixed_patterns(text, N, m, 9, alphabet=acgt)
        _, _, eoff, eocc = g64.locate(patt, N, m)
        d_patt = torch.from_numpy(patt).to(dev)
        d_lo = torch.empty(N + 1, dtype=torch.int64, device=dev); d_hi = torch.empty(N + 1, dtype=torch.int64, device=dev)
        d_off = torch.empty(N + 2, dtype=torch.int64, device=dev)
        cuts, _ = g64.plan_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), 4)
        got = []
        for c0, c1 in zip(cuts, cuts[1:]):
            need = int(eoff[c1] - eoff[c0])
            d_occ = torch.empty(need + 16, dtype=torch.int64, device=dev)
            g64.expand_shard_dev(N, c0, c1, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), d_occ.data_ptr(), need)
            torch.cuda.synchronize()
            got.append(d_occ[:need].cpu().numpy().astype(np.uint64))
        assert np.array_equal(np.concatenate(got), eocc)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
