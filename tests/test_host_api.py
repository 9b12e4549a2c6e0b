"""Host-side pieces of the product that need no GPU: the C-ABI library loads and
exports every symbol include/pm_b200.h declares, the FASTA normaliser and
reverse complement agree with the oracle, and argument errors are reported the
way the header promises.  No scan is run here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(pkg, built_lib):
    hdr = open(os.path.join(ROOT, "include", "pm_b200.h")).read()
    declared = set(re.findall(r"\b(pm_[a-z_]+)\s*\(", hdr))
    declared -= {"pm_ctx"}
    assert declared == set(pkg.EXPORTS)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert b"sm_100a" in built_lib.pm_version()


def test_struct_layouts_match_header(pkg):
    assert C.sizeof(pkg.Hit) == 8 + 8 + 8 + 4 + 4 + 4 + 4 + pkg.PM_MAX_CODES
    assert C.sizeof(pkg.Candidate) == 16
    assert C.sizeof(pkg.PcrOptions) == 6 * 4 + 2 * 8 + 2 * 8 + 8 + 8
    assert C.sizeof(pkg.PcrPair) == 8 + 8 + 8 + 4 + 4


def test_reverse_comp_matches_oracle(pkg, built_lib, oracle):
    for s in ["ACGT", "AAGGGYAATAA", "WKGKWRGGRWW", "acgtnRYKMSWBDHVNX", "U", "TTKGRRGGTAC"]:
        assert pkg.reverse_comp(s) == oracle.reverse_comp(s)


@pytest.mark.parametrize("seed", range(12))
def test_normalize_fasta_matches_oracle(pkg, built_lib, oracle, seed):
    import casegen
    fasta, _ = casegen.make_case(seed, nseq=1 + seed % 5, seqlen=50 + 37 * seed, line=7 + seed,
                                 crlf=bool(seed % 3 == 0), lower=0.05, text_iupac=0.02)
    raw = fasta.encode()
    for uc in (False, True):
        a = pkg.normalize_fasta(raw, uc=uc)
        b = oracle.normalize_fasta(raw, uc=uc)
        assert np.array_equal(a.chars, b.chars)
        assert np.array_equal(a.seq_key, b.seq_key)
        assert a.headers == b.headers


@pytest.mark.parametrize("raw", [b"", b">h\n", b">h\nACGT", b">h\nAC GT\n\n>g\n\nTT\n", b"ACGT\n>x\nGG\n",
                                 b">a\r\nAC\r\nGT\r\n>b\r\nTT\r\n", b">only header"])
def test_normalize_fasta_edge_cases(pkg, built_lib, oracle, raw):
    a = pkg.normalize_fasta(raw)
    b = oracle.normalize_fasta(raw)
    assert np.array_equal(a.chars, b.chars)
    assert np.array_equal(a.seq_key, b.seq_key)


def test_create_rejects_bad_options_or_missing_gpu(pkg, built_lib):
    import torch
    h = C.c_void_p()
    bad = pkg.Options(0, 0, 7, 1, 0, 0, 10)
    assert built_lib.pm_create(C.byref(h), C.byref(bad)) == pkg.PM_E_UNSUPPORTED
    bad = pkg.Options(0, pkg.PM_INDEX_KEYWORD_TREE, 2, 1, 0, 0, 10)
    assert built_lib.pm_create(C.byref(h), C.byref(bad)) == pkg.PM_E_ARG
    if not torch.cuda.is_available():
        ok = pkg.Options(0, 0, 1, 1, 0, 0, 10)
        # no device: the product fails loudly, it does not fall back to a CPU scan
        assert built_lib.pm_create(C.byref(h), C.byref(ok)) == pkg.PM_E_CUDA
        assert b"CUDA" in built_lib.pm_last_error()
        with pytest.raises(pkg.PmError):
            pkg.Scanner(k=1)


def _hits(pkg, rows):
    arr = np.zeros(len(rows), pkg.HIT_DTYPE)
    for i, (pid, start, end, ed) in enumerate(rows):
        arr[i]["pattern_id"], arr[i]["start"], arr[i]["end"], arr[i]["first"], arr[i]["editdist"] = pid, start, end, end, ed
    return arr


def test_pcr_pairs_edge_cases(pkg, built_lib):
    """pm_pcr_pairs is host code: no hits, an odd number of primers (pcr_match.cc:768), hits whose
    own alignment failed, amplicons across two FASTA entries, -m / -M, and the rule that a hit
    only pairs with partners that come after it."""
    keys = np.array([1, 1001], np.int64)           # two entries: [1, 1000) and [1001, ...)
    patlen = [0, 20, 20, 20, 20]                   # one pair, n = 2 (ids 3, 4 = reverse complements)
    assert pkg.pcr_pairs(_hits(pkg, []), 2, patlen, keys) == []
    with pytest.raises(pkg.PmError):
        pkg.pcr_pairs(_hits(pkg, []), 3, patlen + [20, 20], keys)
    fwd, rev = (1, 100, 120, 0), (2, 400, 420, 0)
    assert pkg.pcr_pairs(_hits(pkg, [fwd, rev]), 2, patlen, keys) == [(0, 1, 320, 1)]
    assert pkg.pcr_pairs(_hits(pkg, [rev, fwd]), 2, patlen, keys) == [(1, 0, 320, 1)]     # input order is free
    assert pkg.pcr_pairs(_hits(pkg, [fwd, rev]), 2, patlen, keys, maxdist=300) == []
    assert pkg.pcr_pairs(_hits(pkg, [fwd, rev]), 2, patlen, keys, mindist=321) == []
    assert pkg.pcr_pairs(_hits(pkg, [fwd, rev]), 2, patlen, keys, betweenlen=True, maxdist=300) == [(0, 1, 280, 1)]
    # the reverse primer's site first: nothing (id 2 starts no pair, and it had its turn before id 1 looks)
    assert pkg.pcr_pairs(_hits(pkg, [(2, 100, 120, 0), (1, 400, 420, 0)]), 2, patlen, keys) == []
    # the other strand: id 4 (the reverse primer as given) leads, id 3 follows
    assert pkg.pcr_pairs(_hits(pkg, [(4, 100, 120, 0), (3, 400, 420, 0)]), 2, patlen, keys) == [(0, 1, 320, 1)]
    # a hit with more errors than k is never paired (pcr_match.cc:1114)
    assert pkg.pcr_pairs(_hits(pkg, [fwd, (2, 400, 420, 2)]), 2, patlen, keys, k=1) == []
    assert pkg.pcr_pairs(_hits(pkg, [fwd, (2, 400, 420, -1)]), 2, patlen, keys, k=1) == []
    # an amplicon may not span two entries (fasta_io.t:203-213)
    assert pkg.pcr_pairs(_hits(pkg, [(1, 900, 920, 0), (2, 1100, 1120, 0)]), 2, patlen, keys) == []
    # every later partner inside the window, nearest first; -a adds the same-strand orientations
    got = pkg.pcr_pairs(_hits(pkg, [fwd, rev, (2, 500, 520, 0), (4, 600, 620, 0)]), 2, patlen, keys)
    assert got == [(0, 1, 320, 1), (0, 2, 420, 1)]
    got = pkg.pcr_pairs(_hits(pkg, [fwd, rev, (2, 500, 520, 0), (4, 600, 620, 0)]), 2, patlen, keys, allorient=True)
    assert got == [(0, 1, 320, 1), (0, 2, 420, 1), (0, 3, 520, 1)]


def test_hit_order_replays_the_reference_sort_and_chain_walk(pkg, built_lib, oracle):
    """pm_hit_order (host code of the library: the real std::sort on the candidates' positions, then
    filter_bitvec's chain walk) against the oracle's restatement (std_sort_by_key + the same walk in
    Python) on candidate lists full of equal positions, 1 to 3 000 long."""
    import random
    rng = random.Random(31)
    for trial in range(120):
        k = rng.choice([1, 2, 3])
        npat = rng.choice([1, 2, 5, 16])
        n = rng.choice([1, 5, 16, 17, 40, 200, 3000])
        span = max(4, n // rng.choice([1, 2, 4]))
        cands = sorted({(rng.randrange(1, span + 1), rng.randrange(1, npat + 1)) for _ in range(n)})
        # the oracle's side: candidates as the scan reports them (position, pattern) -> std::sort on the
        # position alone -> chains (same pattern, at most 2k+1 apart) in that order
        srt = oracle.std_sort_by_key(cands, lambda c: c[0])
        done = [False] * len(srt)
        heads = []
        for i, (pos, pid) in enumerate(srt):
            if done[i]:
                continue
            done[i] = True
            last = pos
            heads.append((pid, pos))
            j = i + 1
            while j < len(srt) and (done[j] or srt[j][0] <= last + 2 * k + 1):
                if not done[j] and srt[j][1] == pid:
                    last = srt[j][0]
                    done[j] = True
                j += 1
        # the library's side: one hit per chain (a third of them dropped: chains whose alignment fails
        # report nothing), given in (first, pattern) order as pm_find_patterns returns them
        kept = [h for h in heads if rng.random() < 0.67] or heads[:1]
        given = sorted(kept, key=lambda h: (h[1], h[0]))
        hits = np.zeros(len(given), pkg.HIT_DTYPE)
        for i, (pid, first) in enumerate(given):
            hits[i]["pattern_id"], hits[i]["first"], hits[i]["end"] = pid, first, first + 3
        ca = np.zeros(len(cands), pkg.CAND_DTYPE)
        for i, (pos, pid) in enumerate(sorted(cands, key=lambda c: (c[1], c[0]))):   # as pm_last_candidates sorts them
            ca[i]["pos"], ca[i]["pattern_id"] = pos, pid
        order = np.zeros(len(given), np.uint64)
        rc = built_lib.pm_hit_order(C.c_void_p(hits.ctypes.data), len(hits), C.c_void_p(ca.ctypes.data), len(ca), k,
                                    C.c_void_p(order.ctypes.data))
        assert rc == pkg.PM_OK
        got = [given[int(i)] for i in order]
        want = [h for h in heads if h in set(kept)]
        assert got == want, "trial %d" % trial
