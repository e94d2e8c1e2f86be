"""-m gpu: ri-locate's -o / -c post-processing on the device against the host restatement of
the reference's loop (ri-locate.cpp:146-190: std::sort per pattern; brute-force count + byte comparison)."""
import numpy as np
import pytest

from conftest import rib, ob, mixed_patterns

pytestmark = pytest.mark.gpu


def _sorted_per_pattern(off, occ):
    out = occ.copy()
    for p in range(off.size - 1):
        out[int(off[p]):int(off[p + 1])].sort()
    return out


@pytest.mark.parametrize("variant", ["0", "8"])
def test_sorted_locate_equals_per_pattern_sort(variant, monkeypatch):
    """RIG_LOCATE_SORT == np.sort of every pattern's slice of the plain locate output; every sort tier is hit:
    m = 9 (segments of tens), m = 3 (thousands: shared-memory tiers), m = 1 (tens of thousands and above:
    the large shared-memory tier and the in-place global tier), 32- and 64-bit keys."""
    monkeypatch.setenv("RIG_VARIANT", variant)
    text = rib.gen_text("dna_drift", 600_000, 3_000, 3, 77)
    gpu = rib.GpuIndex(rib.HostIndex.from_text(text))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    tiers = set()
    for (N, m, seed) in [(800, 9, 1), (300, 3, 2), (12, 1, 3), (64, 2, 4)]:
        patt = mixed_patterns(text, N, m, seed, alphabet=acgt)
        lo, hi, off, occ = gpu.locate(patt, N, m)
        lo2, hi2, off2, occ2, rep = gpu.locate_ex(patt, N, m, rib.LOCATE_SORT)
        assert rep is None
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2) and np.array_equal(off, off2)
        assert np.array_equal(occ2, _sorted_per_pattern(off, occ)), "m=%d" % m
        for ln in np.diff(off.astype(np.int64)):
            tiers.add(0 if ln <= 4096 else (1 if ln <= (32768 if variant == "0" else 16384) else 2))
    assert tiers == {0, 1, 2}


def test_device_sort_of_arbitrary_segments():
    """rig_sort_occurrences_dev on synthetic segments: empty, single, power-of-two and odd lengths around the tier caps."""
    torch = pytest.importorskip("torch")
    text = rib.gen_text("dna_drift", 100_000, 1_000, 3, 5)
    gpu = rib.GpuIndex(rib.HostIndex.from_text(text))
    rng = np.random.default_rng(3)
    lens = [0, 1, 2, 3, 31, 32, 33, 1000, 1024, 1025, 4095, 4096, 4097, 5000, 32767, 32768, 32769, 70001, 0, 7]
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(np.array(lens, dtype=np.uint64), out=off[1:])
    vals = rng.integers(0, gpu.n, size=int(off[-1]), dtype=np.uint64)
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_occ = torch.from_numpy(vals.view(np.int64)).cuda()
    gpu.sort_dev(len(lens), d_off.data_ptr(), d_occ.data_ptr(), vals.size, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_occ.cpu().numpy().view(np.uint64), _sorted_per_pattern(off, vals))


def test_check_passes_on_correct_output_and_counts_like_brute_force():
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 11)
    gpu = rib.GpuIndex(rib.HostIndex.from_text(text))
    gpu.text_attach(text)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for (N, m, seed) in [(600, 8, 1), (100, 2, 2), (50, 40, 3)]:
        patt = mixed_patterns(text, N, m, seed, alphabet=acgt)   # includes absent patterns and duplicates
        lo, hi, off, occ, rep = gpu.locate_ex(patt, N, m, rib.LOCATE_SORT | rib.LOCATE_CHECK)
        assert rep.patterns_checked == N and rep.clean, rep.as_dict()
        nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0))
        P = patt.reshape(N, m)
        for p in range(0, N, max(1, N // 25)):
            assert ob.brute_count(text, P[p]) == int(nocc[p])


def test_check_detects_wrong_output():
    """Corrupt the located output on the device and run rig_check_dev: a wrong position, a repeated position,
    a wrong range (count mismatch) and an unsorted slice are each reported."""
    torch = pytest.importorskip("torch")
    text = rib.gen_text("dna_drift", 200_000, 2_000, 3, 13)
    gpu = rib.GpuIndex(rib.HostIndex.from_text(text))
    gpu.text_attach(text)
    N, m = 300, 7
    patt = rib.gen_patterns(text, N, m, 21)
    lo, hi, off, occ, rep = gpu.locate_ex(patt, N, m, rib.LOCATE_SORT)
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    d_patt = torch.from_numpy(patt).to(dev)

    def run(lo_, hi_, occ_, is_sorted=True):
        d_lo = torch.from_numpy(lo_.view(np.int64)).to(dev); d_hi = torch.from_numpy(hi_.view(np.int64)).to(dev)
        d_off = torch.from_numpy(off.view(np.int64)).to(dev); d_occ = torch.from_numpy(occ_.view(np.int64)).to(dev)
        return gpu.check_dev(d_patt.data_ptr(), N, m, d_lo.data_ptr(), d_hi.data_ptr(), d_off.data_ptr(), d_occ.data_ptr(),
                             occ_.size, is_sorted, st)

    assert run(lo, hi, occ).clean
    p = int(np.argmax(np.diff(off.astype(np.int64)) >= 3))
    a = int(off[p])
    bad = occ.copy(); bad[a + 1] = bad[a + 1] + np.uint64(1)            # a position that does not hold the pattern
    r = run(lo, hi, bad)
    assert r.wrong_occurrences >= 1 and r.first_bad_pattern == p
    dup = occ.copy(); dup[a + 1] = dup[a]                                # a repeated position
    r = run(lo, hi, dup)
    assert r.unsorted_or_duplicate >= 1 and r.wrong_occurrences == 0
    hi_bad = hi.copy(); hi_bad[p] = hi_bad[p] + np.uint64(1)             # claims one occurrence more than the text has
    r = run(lo, hi_bad, occ)
    assert r.wrong_count_patterns == 1 and r.first_bad_pattern == p
    rev = occ.copy(); rev[a:int(off[p + 1])] = rev[a:int(off[p + 1])][::-1].copy()
    assert run(lo, hi, rev).unsorted_or_duplicate >= 1
    assert run(lo, hi, rev, is_sorted=False).clean                       # order is only checked on request
    with pytest.raises(rib.RigError):
        rib.GpuIndex(rib.HostIndex.from_text(text)).locate_ex(patt, N, m, rib.LOCATE_CHECK)  # no text attached


def test_edge_cases_sort_and_check():
    text = np.frombuffer(b"abracadabra_abracadabra_\xff\xfe\xffabra" * 5, dtype=np.uint8)
    gpu = rib.GpuIndex(rib.HostIndex.from_text(text))
    with pytest.raises(rib.RigError):
        gpu.text_attach(text[:-1])                                        # not the indexed text
    gpu.text_attach(text)
    lo, hi, off, occ, rep = gpu.locate_ex(np.zeros(0, dtype=np.uint8), 0, 4, rib.LOCATE_SORT | rib.LOCATE_CHECK)
    assert off.tolist() == [0] and rep.clean and rep.patterns_checked == 0
    pats = [b"abra", b"zzzz", b"\xff\xfe\xffa", b"abra", b"a_ab"]
    patt = np.frombuffer(b"".join(pats), dtype=np.uint8)
    lo, hi, off, occ, rep = gpu.locate_ex(patt, len(pats), 4, rib.LOCATE_SORT | rib.LOCATE_CHECK)
    assert rep.clean and np.diff(off.astype(np.int64)).tolist() == [ob.brute_count(text, np.frombuffer(q, dtype=np.uint8)) for q in pats]
    lo, hi, off, occ, rep = gpu.locate_ex(np.zeros(0, dtype=np.uint8), 2, 0, rib.LOCATE_SORT | rib.LOCATE_CHECK)  # m = 0
    assert rep.clean and occ[: text.size + 1].tolist() == list(range(text.size + 1))


def test_locate32_equals_locate():
    """rig_locate_batch32: the same positions as rig_locate_batch, as uint32, plain and sorted; refused when
    positions would not fit (64-bit index)."""
    text = rib.gen_text("dna_drift", 300_000, 3_000, 3, 19)
    host = rib.HostIndex.from_text(text)
    gpu = rib.GpuIndex(host)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for (N, m, seed) in [(700, 8, 1), (0, 5, 2), (33, 1, 3), (17001, 11, 4)]:
        patt = mixed_patterns(text, N, m, seed, alphabet=acgt) if N else np.zeros(0, dtype=np.uint8)
        lo, hi, off, occ = gpu.locate(patt, N, m)
        lo2, hi2, off2, occ32 = gpu.locate32(patt, N, m)
        assert occ32.dtype == np.uint32
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2) and np.array_equal(off, off2)
        assert np.array_equal(occ32.astype(np.uint64), occ)
        _, _, _, occ32s = gpu.locate32(patt, N, m, rib.LOCATE_SORT)
        assert np.array_equal(occ32s.astype(np.uint64), _sorted_per_pattern(off, occ))
    with pytest.raises(rib.RigError):
        gpu.locate32(patt, N, m, rib.LOCATE_CHECK)
    # native 32-bit output of the expansion kernels (default table form: fused, small windows, single pass) and the
    # narrowing pass behind the other forms: long ranges, so that chains span many windows and lines
    for (jump, seg) in [(0, 0), (4, 16), (4, 1), (2, 0), (8, 32)]:
        g2 = rib.GpuIndex(host, phi_jump=jump, seed_jump=seg)
        for (N, m, seed) in [(300, 3, 7), (41, 1, 8), (5000, 6, 9)]:
            patt = mixed_patterns(text, N, m, seed, alphabet=acgt)
            lo, hi, off, occ = g2.locate(patt, N, m)
            lo2, hi2, off2, occ32 = g2.locate32(patt, N, m)
            assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2) and np.array_equal(off, off2)
            assert np.array_equal(occ32.astype(np.uint64), occ), (jump, seg, N, m)
