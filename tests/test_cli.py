"""CLI contract: ri-build / ri-count / ri-locate keep the reference's usage text, options, exit codes
and stdout lines (reference ri-build.cpp, ri-count.cpp, ri-locate.cpp). Where oracle/_ref holds the
reference's own binaries (its unmodified mains over the SDSL-API shim) the outputs are diffed
against them; timing lines are masked. CPU tier: everything that needs no query; GPU tier: the rest."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import rib, ob, ROOT, needs_ref

BIN = os.path.join(ROOT, "r-index_b200", "bin")
REFBIN = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module", autouse=True)
def _built():
    rib.build_gpu()
    rib.build_cli()


def run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, errors="replace", **kw)


def mask(s):
    s = re.sub(r"(Load time|Total time|Build time) : .*", r"\1 : <t>", s)
    s = re.sub(r"Search time : \S+ (milliseconds/\w+)", r"Search time : <t> \1", s)
    s = "\n".join(l for l in s.splitlines() if not l.startswith("[gpu]"))
    return s


def make_inputs(tmp_path, n=60_000, N=300, m=7):
    text = rib.gen_text("dna_drift", n, 600, 2, 77)
    tfile = str(tmp_path / "text.txt")
    open(tfile, "wb").write(bytes(text))
    patt = rib.gen_patterns(text, N, m, 5)
    patt[:m] = np.frombuffer(b"\nA\xffC\nGT"[:m], dtype=np.uint8)  # binary-safe: newline / >=0x80 bytes in a pattern
    pfile = str(tmp_path / "p.patt")
    rib.write_pattern_file(pfile, patt, N, m, "text.txt")
    return tfile, pfile, text, patt


@needs_ref
def test_usage_texts_match_reference():
    for tool in ("ri-build", "ri-count", "ri-locate"):
        ours, ref = run([os.path.join(BIN, tool)]), run([os.path.join(REFBIN, tool)])
        assert ours.returncode == ref.returncode == 0
        assert ours.stdout == ref.stdout
    ours = run([os.path.join(BIN, "ri-locate"), "-x", "a", "b"])
    ref = run([os.path.join(REFBIN, "ri-locate"), "-x", "a", "b"])
    assert ours.stdout == ref.stdout and ours.stdout.startswith("Error: unknown option -x")
    ours = run([os.path.join(BIN, "ri-build"), "-z", "f"])
    ref = run([os.path.join(REFBIN, "ri-build"), "-z", "f"])
    assert ours.stdout == ref.stdout and "unrecognized '-z' option" in ours.stdout


@needs_ref
def test_ri_build_stdout_and_file_framing(tmp_path):
    tfile, _, text, _ = make_inputs(tmp_path)
    ours = run([os.path.join(BIN, "ri-build"), "-o", str(tmp_path / "ours"), tfile])
    ref = run([os.path.join(REFBIN, "ri-build"), "-o", str(tmp_path / "ref"), tfile])
    assert ours.returncode == ref.returncode == 0
    sub = lambda s: mask(s).replace(str(tmp_path / "ours"), "X").replace(str(tmp_path / "ref"), "X")  # noqa: E731
    assert sub(ours.stdout) == sub(ref.stdout)       # n, r, n/r, log2 lines included
    blob = open(str(tmp_path / "ours.ri"), "rb").read()
    assert blob[0] == 0 and blob[1:9] == b"RIB200v1"   # `fast` flag byte first (ri-build.cpp:133)
    h = rib.HostIndex.load(str(tmp_path / "ours.ri"))
    assert h.n == text.size + 1


def test_reserved_bytes_exit_code(tmp_path):
    f = str(tmp_path / "bad.txt")
    open(f, "wb").write(b"abc\x01def")
    out = run([os.path.join(BIN, "ri-build"), f])
    assert out.returncode == 1
    assert "Error: input string contains one of the reserved characters 0x0, 0x1" in out.stdout


def test_malformed_pattern_header_exits_zero(tmp_path):
    """utils.hpp:51-55: message + exit(0). Checked before any GPU work."""
    tfile, _, _, _ = make_inputs(tmp_path, n=5000, N=5, m=3)
    assert run([os.path.join(BIN, "ri-build"), tfile]).returncode == 0
    bad = str(tmp_path / "bad.patt")
    open(bad, "wb").write(b"# nombre=3 length=2 file=x forbidden=\nabcdef")
    for tool in ("ri-count", "ri-locate"):
        out = run([os.path.join(BIN, tool), tfile + ".ri", bad])
        assert out.returncode == 0 and "Error: malformed header in patterns file" in out.stdout
        assert out.stdout.startswith("Loading r-index\nsearching patterns ... \nError: malformed header")


@pytest.mark.gpu
@needs_ref
def test_ri_count_and_locate_match_reference_cli(tmp_path):
    tfile, pfile, text, patt = make_inputs(tmp_path)
    assert run([os.path.join(BIN, "ri-build"), "-o", str(tmp_path / "ours"), tfile]).returncode == 0
    assert run([os.path.join(REFBIN, "ri-build"), "-o", str(tmp_path / "ref"), tfile]).returncode == 0
    ours = run([os.path.join(BIN, "ri-count"), str(tmp_path / "ours.ri"), pfile])
    ref = run([os.path.join(REFBIN, "ri-count"), str(tmp_path / "ref.ri"), pfile])
    assert ours.returncode == 0, ours.stdout + ours.stderr
    assert mask(ours.stdout) == mask(ref.stdout)
    assert "total number of occurrences  occ_t = " in ours.stdout   # the double space is upstream's
    o1, o2 = str(tmp_path / "o1.txt"), str(tmp_path / "o2.txt")
    ours = run([os.path.join(BIN, "ri-locate"), "-c", tfile, "-o", o1, str(tmp_path / "ours.ri"), pfile])
    ref = run([os.path.join(REFBIN, "ri-locate"), "-c", tfile, "-o", o2, str(tmp_path / "ref.ri"), pfile])
    assert ours.returncode == 0, ours.stdout + ours.stderr
    assert mask(ours.stdout) == mask(ref.stdout)
    assert "Error" not in ours.stdout
    assert open(o1, "rb").read() == open(o2, "rb").read()          # -o file byte-identical
    malformed = str(tmp_path / "bad.patt")
    open(malformed, "wb").write(b"# number=3 length=2\nabcdef")
    out = run([os.path.join(BIN, "ri-count"), str(tmp_path / "ours.ri"), malformed])
    assert out.returncode == 0 and "Error: malformed header in patterns file" in out.stdout


@pytest.mark.gpu
def test_cli_multi_gpu_flag_is_shard_invariant(tmp_path):
    """--gpus G (clamped to the devices present) must not change any printed result."""
    tfile, pfile, _, _ = make_inputs(tmp_path, n=40_000, N=257, m=6)
    assert run([os.path.join(BIN, "ri-build"), tfile]).returncode == 0
    a = run([os.path.join(BIN, "ri-locate"), "-o", str(tmp_path / "a"), tfile + ".ri", pfile])
    b = run([os.path.join(BIN, "ri-locate"), "--gpus", "8", "-o", str(tmp_path / "b"), tfile + ".ri", pfile])
    assert a.returncode == 0 and b.returncode == 0
    assert mask(a.stdout) == mask(b.stdout)
    assert open(str(tmp_path / "a"), "rb").read() == open(str(tmp_path / "b"), "rb").read()


@pytest.mark.gpu
def test_cli_flat_file_is_written_once_and_reused(tmp_path):
    """--flat F: the first run flattens and writes F, later runs (ri-locate and ri-count alike) load it instead of
    flattening; no printed result changes; a flat file of another index is ignored (the index is flattened)."""
    tfile, pfile, _, _ = make_inputs(tmp_path, n=60_000, N=300, m=7)
    assert run([os.path.join(BIN, "ri-build"), tfile]).returncode == 0
    flat = str(tmp_path / "idx.flat")
    plain = run([os.path.join(BIN, "ri-locate"), "-c", tfile, "-o", str(tmp_path / "p"), tfile + ".ri", pfile])
    first = run([os.path.join(BIN, "ri-locate"), "--flat", flat, "-c", tfile, "-o", str(tmp_path / "a"), tfile + ".ri", pfile])
    assert first.returncode == 0 and os.path.getsize(flat) > 0
    again = run([os.path.join(BIN, "ri-locate"), "--flat", flat, "-c", tfile, "-o", str(tmp_path / "b"), tfile + ".ri", pfile])
    assert again.returncode == 0 and "Error" not in again.stdout
    assert mask(plain.stdout) == mask(first.stdout) == mask(again.stdout)
    assert open(str(tmp_path / "p"), "rb").read() == open(str(tmp_path / "a"), "rb").read() == open(str(tmp_path / "b"), "rb").read()
    c0 = run([os.path.join(BIN, "ri-count"), tfile + ".ri", pfile])
    c1 = run([os.path.join(BIN, "ri-count"), "--flat", flat, tfile + ".ri", pfile])
    assert c0.returncode == 0 and c1.returncode == 0 and mask(c0.stdout) == mask(c1.stdout)
    (tmp_path / "other").mkdir()
    t2, p2, _, _ = make_inputs(tmp_path / "other", n=30_000, N=50, m=5)
    assert run([os.path.join(BIN, "ri-build"), t2]).returncode == 0
    o0 = run([os.path.join(BIN, "ri-count"), t2 + ".ri", p2])
    o1 = run([os.path.join(BIN, "ri-count"), "--flat", flat, t2 + ".ri", p2])   # not this index's file
    assert o1.returncode == 0 and mask(o0.stdout) == mask(o1.stdout)


def test_ri_build_large_text_takes_the_pfp_route_and_equals_sais(tmp_path):
    """From 16 MB up ri-build constructs by prefix-free parsing; the .ri it writes holds exactly
    the arrays the in-memory SA-IS route gives (forced with RIB_BUILDER=sais), and the stdout lines are the same."""
    text = rib.gen_text("dna_drift", 17_000_000, 20_000, 3, 5)
    tfile = str(tmp_path / "big.txt")
    open(tfile, "wb").write(bytes(text))
    a = run([os.path.join(BIN, "ri-build"), "-o", str(tmp_path / "pfp"), tfile])
    b = run([os.path.join(BIN, "ri-build"), "-o", str(tmp_path / "sais"), tfile], env=dict(os.environ, RIB_BUILDER="sais"))
    assert a.returncode == 0 and b.returncode == 0
    assert mask(a.stdout).replace("pfp.ri", "X") == mask(b.stdout).replace("sais.ri", "X")
    A, B = rib.HostIndex.load(str(tmp_path / "pfp.ri")).arrays(), rib.HostIndex.load(str(tmp_path / "sais.ri")).arrays()
    assert (A["n"], A["r"]) == (B["n"], B["r"]) == (text.size + 1, B["r"])
    for k in ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"):
        assert np.array_equal(A[k], B[k]), k
    assert rib.HostIndex.from_text_auto(text).used_pfp is True
