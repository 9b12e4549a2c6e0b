"""bench.py's parity leg on the CPU: the reference runs on windows cut out of a shard and its
records are mapped back to stream coordinates; here the oracle stands in for the GPU hit list
(this checks the window / coordinate logic of bench.py, not the kernels)."""
import importlib
import os
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "primer_match")),
                    reason="needs the compiled reference (oracle/_ref)")
def test_parity_windows_maps_reference_records_to_stream_coordinates(pkg, oracle, monkeypatch):
    import torch
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "SEQ_LEN", 150_000)
    cfg = dict(bench.CONFIGS["headline"], nprimers=30)
    primers = bench.make_primers(cfg)
    g = np.random.default_rng(5)
    n = 450_000
    T = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, n + 1)].copy()
    T[0::150_000] = 10
    T[n] = 10
    pats = oracle.pattern_array(primers, True)
    planted = []
    for j in range(40):
        p = bytearray(pats[j % len(pats)].encode())
        for _ in range(j % 3):
            p[int(g.integers(0, 20))] = b"ACGT"[int(g.integers(0, 4))]
        at = int(g.integers(1000, n - 1000))
        if at % 150_000 < 100 or at % 150_000 > 149_900:
            continue
        T[at:at + 20] = np.frombuffer(bytes(p), np.uint8)
        planted.append(at)
    keys = np.array([i + 1 for i in range(0, n, 150_000)], np.int64)
    st = oracle.Stream(T, keys, ["s%d" % i for i in range(len(keys))])
    recs = oracle.primer_match(st, primers, k=2, indels=False, rev_comp=True)
    hits = np.zeros(len(recs), pkg.HIT_DTYPE)
    for i, r in enumerate(recs):
        hits[i]["start"], hits[i]["end"] = r.stream_start, r.stream_end
        hits[i]["pattern_id"] = r.pattern_index + (len(primers) if r.strand == "R" else 0)
        hits[i]["editdist"] = r.editdist
    assert len(hits) >= 30
    work = tempfile.mkdtemp(prefix="pmwin_")
    dt, bp, compared, ok, detail = bench.parity_windows(cfg, primers, torch.from_numpy(T), n + 1, hits, 4, 60_000, work, planted)
    assert ok, detail
    assert compared >= 5 and bp == 4 * 60_000
    # a corrupted hit list must be noticed
    bad = hits.copy()
    inside = [i for i in range(len(bad)) if any(a <= bad[i]["start"] - 300 and bad[i]["end"] + 300 <= a + 60_000
                                                for a in [max(1, 150_000 - 30_000), 1])]
    if inside:
        bad[inside[0]]["editdist"] += 1
        _, _, _, ok2, _ = bench.parity_windows(cfg, primers, torch.from_numpy(T), n + 1, bad, 4, 60_000, work, planted)
        assert not ok2
