#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE'S OWN CODE.

Run in the dev container (needs /root/reference to build oracle/_ref). The fixtures pin the
oracle port (tests/test_oracle.py::test_golden_fixtures) and the CUDA path
(tests/test_gpu_golden.py) to what the reference computes: ranges from r_index<>::count,
occurrences from r_index<>::locate_all (order included), and the logical index content read
through the reference's accessors (ref_extract in oracle/ref_driver.cpp).

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

rib = ge.load_package()
ob = ge.load_oracle()
from conftest import mixed_patterns  # noqa: E402


def make(name, text, N, m, seed, alphabet=None):
    ref = ob.RefIndex.from_text(text)
    patt = mixed_patterns(text, N, m, seed, alphabet=alphabet)
    lo, hi, off, occ, _ = ref.locate(patt, N, m)
    ex = ref.extract()
    idx_sha = hashlib.sha256(b"".join(np.ascontiguousarray(ex[k]).tobytes() for k in
                                      ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"))).hexdigest()
    out = dict(text=text, patterns=patt, N=N, m=m, lo=lo, hi=hi, occ_offsets=off,
               occ_sha256=hashlib.sha256(occ.tobytes()).hexdigest(), index_sha256=idx_sha, r=ex["r"], n=ex["n"])
    if occ.size <= 60_000:
        out["occ"] = occ
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-18s n=%d r=%d N=%d m=%d occ=%d" % (name, ex["n"], ex["r"], N, m, occ.size))


def main():
    ob.build()
    assert ob.have_ref(), "oracle/_ref is needed (build it where /root/reference exists)"
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    make("dna_drift_40k", rib.gen_text("dna_drift", 40_000, 800, 2, 0x601D), 300, 8, 1, acgt)
    make("dna_indep_60k", rib.gen_text("dna_indep", 60_000, 1_500, 2_000_000, 0x601E), 200, 12, 2, acgt)
    make("doc_sigma96_30k", rib.gen_text("versioned_doc", 30_000, 600, 96, 0x601F), 250, 6, 3)
    make("pangenome_50k", rib.gen_text("pangenome", 50_000, 1_000, 40, 0x6020), 200, 10, 4,
         np.frombuffer(b"ACGTN", dtype=np.uint8))
    edge = np.frombuffer((b"abracadabra_\xff\xfe\xff_abracadabra\n" * 40) + b"\xff", dtype=np.uint8)
    make("edge_bytes", edge, 120, 3, 5)
    make("single_symbol", np.frombuffer(b"a" * 500, dtype=np.uint8), 20, 4, 6)


if __name__ == "__main__":
    main()
