#!/usr/bin/env python
"""Config 1 (BASELINE.json configs[0]): the reference's bundled world_leaders text + its
Pizza&Chili pattern file. Run in the dev container (needs /root/reference/datasets and oracle/_ref):

  * decodes the two members with tools/sevenz.py (CRC-checked),
  * runs the REFERENCE code (oracle/_ref) for count + locate_all over the 1000 patterns,
  * writes tests/golden/config1_world_leaders.json (counts digest, occurrence digests, occ_t) — small,
    committed — and caches the index built by this repo's builder + the pattern file under .cache/
    (not committed; travels to the GPU box with the snapshot) for tests/test_gpu_config1.py.

The known answer occ_t = 29,781,174 (brute force over the text) is asserted here too.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402
import sevenz  # noqa: E402

rib = ge.load_package()
ob = ge.load_oracle()
from rindex_b200._gpu import digest_host  # noqa: E402


def main():
    ds = "/root/reference/datasets"
    text = np.frombuffer(sevenz.extract_one(os.path.join(ds, "texts.7z"), "world_leaders"), dtype=np.uint8)
    pfile = sevenz.extract_one(os.path.join(ds, "patterns.7z"), "world_leaders_1000_8.patt")
    N, m, patt = rib.parse_pattern_file(pfile)
    ref = ob.RefIndex.from_text(text)
    lo, hi, off, occ, _ = ref.locate(patt, N, m, threads=8)
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
    assert int(nocc.sum()) == 29_781_174 and occ.size == 29_781_174
    assert hashlib.sha256(nocc.astype("<u8").tobytes()).hexdigest()[:16] == "f5e5ac89a564dacc"
    s0, s1 = digest_host(occ)
    out = {"dataset": "world_leaders", "text_bytes": int(text.size), "N": N, "m": m, "n": int(ref.n), "r": int(ref.r),
           "occ_t": int(nocc.sum()), "counts_sha256": hashlib.sha256(nocc.astype("<u8").tobytes()).hexdigest(),
           "lo_sha256": hashlib.sha256(lo.tobytes()).hexdigest(), "hi_sha256": hashlib.sha256(hi.tobytes()).hexdigest(),
           "occ_sha256": hashlib.sha256(occ.tobytes()).hexdigest(), "occ_digest": [s0, s1],
           "source": "reference code (oracle/_ref) run by tests/golden/make_config1.py"}
    json.dump(out, open(os.path.join(HERE, "config1_world_leaders.json"), "w"), indent=1)
    os.makedirs(os.path.join(ROOT, ".cache"), exist_ok=True)
    host = rib.HostIndex.from_text(text)
    a, b = host.arrays(), ref.extract()
    for k in ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k   # own builder == reference-built index
    host.save(os.path.join(ROOT, ".cache", "world_leaders.rib"))
    open(os.path.join(ROOT, ".cache", "world_leaders_1000_8.patt"), "wb").write(pfile)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
