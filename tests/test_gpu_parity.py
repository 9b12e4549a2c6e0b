"""Parity of the CUDA path (through the C ABI) with the oracle and with the
reference's golden vectors.  Integer / index work: the bar is bit-exact."""
import random

import numpy as np
import pytest

from conftest import as_tuples, records_of

pytestmark = pytest.mark.gpu

TILE_POS = 32768


def _n5(pkg, params):
    """The goldens of make_golden.py and the oracle's default are `-N 5` (filter_bitvec) for k > 0."""
    return pkg.PM_INDEX_FILTER_BITVEC if params["k"] > 0 else pkg.PM_INDEX_AUTO


def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device; the product has no CPU fallback")


@pytest.fixture(scope="module", autouse=True)
def _gpu(built_lib):
    _need_gpu()


@pytest.mark.parametrize("path", ["auto", "planes"])
def test_golden_records_bit_exact(pkg, golden, path):
    """auto = pair-seed index kernels where they apply (DNA / IUPAC primers whose
    pieces are long enough), bit-plane kernels otherwise; planes = bit-plane
    kernels everywhere."""
    sp = pkg.PM_SCAN_AUTO if path == "auto" else pkg.PM_SCAN_PLANES
    for case in golden["cases"]:
        st = pkg.normalize_fasta(case["fasta"].encode("latin-1"))
        got = pkg.primer_match(st, case["primers"], scan_path=sp, index=_n5(pkg, case["params"]), **case["params"])
        assert as_tuples(got) == sorted(records_of(case)), "case %s" % case["seed"]


@pytest.mark.parametrize("path", ["auto", "planes"])
def test_exact_halves_golden_records_bit_exact(pkg, halves_golden, path):
    """k = 1 with the automatic index choice (exact_halves, select.cc:120-126) and exact_halves
    forced with k = 2: the records the reference printed."""
    sp = pkg.PM_SCAN_AUTO if path == "auto" else pkg.PM_SCAN_PLANES
    for case in halves_golden["cases"]:
        st = pkg.normalize_fasta(case["fasta"].encode("latin-1"))
        index = pkg.PM_INDEX_AUTO if case["select"] == 0 else pkg.PM_INDEX_EXACT_HALVES
        got = pkg.primer_match(st, case["primers"], scan_path=sp, index=index, **case["params"])
        assert as_tuples(got) == sorted(records_of(case)), "case %s" % case["seed"]


def test_pcr_match_golden_records_bit_exact(pkg, oracle, pcr_golden):
    """pcr_match: scan + alignments on the GPU (pm_find_patterns), pairing by pm_pcr_pairs; every
    record the reference's pcr_match printed for its 80 command lines, in its order (-a, -b,
    -m / -M, constraints, -N 0 / 5, UniSTS sizes with -d)."""
    import dataclasses
    n = 0
    for case in pcr_golden["cases"]:
        st = pkg.normalize_fasta(case["fasta"].encode("latin-1"))
        par = dict(case["params"])
        select = par.pop("select")
        got = pkg.pcr_match(st, case["primers"], index=pkg.PM_INDEX_AUTO if select == 0 else pkg.PM_INDEX_FILTER_BITVEC,
                            sts_sizes=case.get("sts_sizes"), **par)
        want = oracle.parse_pcr_records(case["stdout_records"])
        assert [dataclasses.astuple(r) for r in got] == [dataclasses.astuple(r) for r in want], "case %s" % case["seed"]
        n += len(got)
    assert n > 300


def test_pcr_match_orders_simultaneous_hits_like_the_reference(pkg, oracle, pcr_ties_golden):
    """Hits of different primers that end at one position: pm_last_candidates + pm_pcr_pairs
    reproduce the order the reference's two std::sort calls leave them in (and with it which of
    two partners finds the other).  The records the reference printed, in its order."""
    import dataclasses
    n = 0
    for case in pcr_ties_golden["cases"]:
        st = pkg.normalize_fasta(case["fasta"].encode("latin-1"))
        par = dict(case["params"])
        select = par.pop("select")
        got = pkg.pcr_match(st, case["primers"], index=pkg.PM_INDEX_AUTO if select == 0 else pkg.PM_INDEX_FILTER_BITVEC,
                            **par)
        want = oracle.parse_pcr_records(case["stdout_records"])
        assert [dataclasses.astuple(r) for r in got] == [dataclasses.astuple(r) for r in want], "case %s" % case["seed"]
        n += len(got)
    assert n > 400


@pytest.mark.parametrize("path", ["auto", "planes"])
def test_constraints_golden_records_bit_exact(pkg, constraints_golden, path):
    """-s / -e / -5 / -3 (exact_start_bases / exact_end_bases) with -N 5 and with the automatic
    index choice: exact_bases from 6 constrained characters on, exact_halves / filter_bitvec
    below.  The records the reference printed."""
    sp = pkg.PM_SCAN_AUTO if path == "auto" else pkg.PM_SCAN_PLANES
    for case in constraints_golden["cases"]:
        st = pkg.normalize_fasta(case["fasta"].encode("latin-1"))
        index = pkg.PM_INDEX_AUTO if case["select"] == 0 else pkg.PM_INDEX_FILTER_BITVEC
        got = pkg.primer_match(st, case["primers"], scan_path=sp, index=index, **case["params"],
                               **case["constraints"])
        assert as_tuples(got) == sorted(records_of(case)), "case %s (%s)" % (case["seed"], case["index"])


@pytest.mark.parametrize("seed", range(3000, 3040))
def test_exact_halves_matches_oracle_seeded(pkg, oracle, seed):
    import casegen
    params, kw = _params(seed)
    params["k"] = kw["k"] = 1
    fasta, primers = casegen.make_case(seed, **kw)
    if min(len(p) for p in primers) < 8:
        pytest.skip("pick_pattern_index takes filter_bitvec for primers shorter than 8")
    raw = fasta.encode()
    want = oracle.primer_match(oracle.normalize_fasta(raw), primers, index="exact_halves", **params)
    got = pkg.primer_match(pkg.normalize_fasta(raw), primers, **params)      # PM_INDEX_AUTO
    assert as_tuples(got) == as_tuples(want)


@pytest.mark.parametrize("seed", range(4000, 4024))
def test_long_primers_match_oracle(pkg, oracle, seed):
    """Primers of 33-64 characters (bit-plane scan; the halves / exact parts of exact_halves and
    exact_bases still go through the pair index), every index selection, with and without
    exact-base constraints."""
    import casegen
    rng = random.Random(seed)
    k = rng.choice([1, 1, 2, 3])
    indels = rng.random() < 0.5
    wc = rng.random() < 0.3
    params = dict(k=k, indels=indels, wc=wc, tn=wc and rng.random() < 0.5, rev_comp=rng.random() < 0.7)
    kw = dict(nseq=rng.choice([1, 3]), seqlen=rng.choice([400, 2500]), nprimers=rng.choice([2, 6]),
              plen=rng.choice([(33, 40), (40, 64), (64, 64)]), k=k, indels=indels, wc=wc,
              text_iupac=rng.choice([0, 0.01]), lower=0, repeats=rng.random() < 0.5, line=60, crlf=False)
    fasta, primers = casegen.make_case(seed, **kw)
    cons = dict(stlen=0, edlen=0, fplen=rng.choice([0, 0, 8, 20]), tplen=rng.choice([0, 0, 6, 12]))
    raw = fasta.encode()
    for index, oname in ((pkg.PM_INDEX_AUTO, "auto"), (pkg.PM_INDEX_FILTER_BITVEC, "filter_bitvec")):
        want = oracle.primer_match(oracle.normalize_fasta(raw), primers, index=oname, **params, **cons)
        got = pkg.primer_match(pkg.normalize_fasta(raw), primers, index=index, **params, **cons)
        assert as_tuples(got) == as_tuples(want), oname


def _params(seed):
    rng = random.Random(seed)
    k = rng.choice([0, 1, 2, 3])
    indels = rng.random() < 0.5
    wc = rng.random() < 0.4
    tn = wc and rng.random() < 0.5
    rc = rng.random() < 0.6
    kw = dict(nseq=rng.choice([1, 3, 5]), seqlen=rng.choice([60, 400, 1500]),
              nprimers=rng.choice([1, 2, 6, 12]), plen=rng.choice([(8, 12), (12, 22), (20, 20), (25, 32)]),
              k=k, indels=indels, wc=wc, text_iupac=rng.choice([0, 0, 0.01, 0.05]),
              lower=0.0 if (wc and k == 0) else rng.choice([0, 0, 0.02]), repeats=rng.random() < 0.7,
              line=rng.choice([60, 70, 13]), crlf=rng.random() < 0.15)
    return dict(k=k, indels=indels, wc=wc, tn=tn, rev_comp=rc), kw


@pytest.mark.parametrize("seed", range(2000, 2080))
def test_records_match_oracle_seeded(pkg, oracle, seed):
    import casegen
    params, kw = _params(seed)
    fasta, primers = casegen.make_case(seed, **kw)
    raw = fasta.encode()
    want = oracle.primer_match(oracle.normalize_fasta(raw), primers, **params)
    for sp in (pkg.PM_SCAN_AUTO, pkg.PM_SCAN_PLANES):
        got = pkg.primer_match(pkg.normalize_fasta(raw), primers, scan_path=sp, index=_n5(pkg, params), **params)
        assert as_tuples(got) == as_tuples(want), "scan path %d" % sp


@pytest.mark.parametrize("path", ["index", "planes"])
@pytest.mark.parametrize("k,indels,wc", [(0, False, False), (0, False, True), (1, False, False),
                                         (2, False, False), (3, False, True), (1, True, False),
                                         (2, True, False), (2, True, True), (3, True, False)])
def test_candidates_match_oracle_across_tiles(pkg, oracle, k, indels, wc, path):
    """Raw candidate set of the dense scan on a text of several tiles, with
    occurrences planted across the tile seams and the bit-row seams."""
    import casegen
    rng = random.Random(100 * k + 10 * indels + wc)
    n = 2 * TILE_POS + 5000
    # the pair-seed index needs k+2 pieces of at least 4 characters from every primer
    minlen = 4 * (k + 2) if path == "index" else max(8, k + 6)
    primers = [casegen.rand_dna(rng, rng.randint(minlen, 32)) for _ in range(10)]
    primers[0] = "A" * max(14, minlen)
    if wc:
        primers[1] = "ACGRYACGTNACGTKMAC"
        primers[2] = "WSACGTBDHVACGT"
        for i in (1, 2):
            primers[i] += casegen.rand_dna(rng, max(0, minlen - len(primers[i])))
    seq = list(casegen.rand_dna(rng, n))
    seams = [1024 * j + d for j in (1, 2, 31, 32, 33, 63, 64) for d in (-20, -3, 0, 5)] + \
            [TILE_POS - 10, TILE_POS - 1, TILE_POS, 2 * TILE_POS - 7, n - 25, 3]
    for at in seams:
        p = rng.choice(primers)
        q = "".join(c if c in "ACGT" else rng.choice("ACGT") for c in p)
        q = casegen.mutate(rng, q, rng.randint(0, k), indels)
        at = max(0, min(n - len(q), at))
        seq[at:at + len(q)] = list(q)
    seq[700:760] = list("A" * 60)
    seq[TILE_POS - 30:TILE_POS + 30] = list("A" * 60)
    for i in (5000, 5001, 40000):
        seq[i] = "N"
    if wc:   # ambiguous text characters, some of them inside planted occurrences
        for at in seams[::3]:
            seq[max(0, min(n - 1, at + rng.randint(0, 12)))] = rng.choice("RYKMSWBDHVN")
        for _ in range(150):
            seq[rng.randrange(n)] = rng.choice("RYKMSWBDHVN")
    seq = "".join(seq)
    cut = 40123
    fasta = ">a\n%s\n>b\n%s\n" % (seq[:cut], seq[cut:])
    st = pkg.normalize_fasta(fasta.encode())
    T = st.chars
    if k == 0:
        want = oracle.exact_scan(T, primers, wc, False)
    else:
        want = oracle.wm_scan(T, primers, k, indels, wc, False)
    want = sorted({(int(b), int(a)) for a, b, _ in want})
    sc = pkg.Scanner(k=k, indels=indels, wildcards=wc, index=_n5(pkg, dict(k=k)),
                     scan_path=pkg.PM_SCAN_INDEX if path == "index" else pkg.PM_SCAN_PLANES)
    for i, p in enumerate(primers):
        sc.add_pattern(p, i + 1)
    sc.init()      # PM_SCAN_INDEX on patterns it cannot serve fails here: never skipped
    sc.set_text(T)
    got = sc.scan_candidates()
    assert sc.stats()["scan_path"] == (pkg.PM_SCAN_INDEX if path == "index" else pkg.PM_SCAN_PLANES)
    got = sorted((int(b), int(a)) for a, b in got)
    sc.close()
    assert len(want) > 20
    assert got == want


@pytest.mark.parametrize("path", ["auto", "planes"])
@pytest.mark.parametrize("k,indels", [(1, False), (2, True)])
def test_shards_reproduce_whole_scan(pkg, k, indels, path):
    """Partitioning the stream over several contexts (GPUs) with the documented
    overlap gives exactly the hits of one whole-stream scan."""
    import casegen
    rng = random.Random(77 + k)
    primers = [casegen.rand_dna(rng, 20) for _ in range(8)] + ["A" * 16]
    n = 3 * TILE_POS
    seq = list(casegen.rand_dna(rng, n))
    for _ in range(120):
        q = casegen.mutate(rng, rng.choice(primers), rng.randint(0, k), indels)
        at = rng.randrange(0, n - 40)
        seq[at:at + len(q)] = list(q)
    cuts = [30000, 61234, 80000]
    for c in cuts:   # occurrences and a poly-A run straddling every cut
        seq[c - 8:c + 12] = list(primers[0])
    seq[cuts[1] + 100:cuts[1] + 160] = list("A" * 60)
    st = pkg.normalize_fasta((">s\n" + "".join(seq) + "\n").encode())
    T = st.chars

    def run(begin, end, lo, hi):
        sc = pkg.Scanner(k=k, indels=indels, index=pkg.PM_INDEX_FILTER_BITVEC,
                         scan_path=pkg.PM_SCAN_AUTO if path == "auto" else pkg.PM_SCAN_PLANES)
        for i, p in enumerate(primers):
            sc.add_pattern(p, i + 1)
        sc.init()
        ov = sc.shard_overlap()
        a = max(0, lo - ov) if lo else 0
        sc.set_text(T[a:hi], offset=a)
        h = sc.find_patterns(begin, end)
        sc.close()
        return h

    whole = run(0, pkg.INT64_MAX, 0, len(T))
    parts = []
    bounds = [0] + cuts + [len(T) + 1]
    for i in range(len(bounds) - 1):
        hi = min(len(T), bounds[i + 1] + 70000)
        parts.append(run(bounds[i], bounds[i + 1], bounds[i], hi))
    merged = np.concatenate(parts)
    merged = merged[np.lexsort((merged["pattern_id"], merged["first"]))]
    assert len(whole) > 50
    assert merged.tobytes() == whole.tobytes()


@pytest.mark.parametrize("k,indels,halves", [(0, False, False), (1, False, False), (1, True, False),
                                             (1, False, True), (1, True, True)])
def test_large_pattern_set_uses_the_l2_filter(pkg, k, indels, halves):
    """52 000 primers x 2 strands: the 2^20-bit level-1 bitmap saturates and the L2-resident
    level-1b table takes over.  Same hits as the bit-plane kernels."""
    import casegen
    rng = random.Random(40 + k)
    primers = [casegen.rand_dna(rng, 20) for _ in range(52000)]
    n = 500_000
    g = np.random.default_rng(8)
    seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, n)].copy()
    for j in range(300):
        q = casegen.mutate(rng, primers[rng.randrange(len(primers))], rng.randint(0, k), indels)
        at = 500 + j * 1600
        seq[at:at + len(q)] = np.frombuffer(q.encode(), np.uint8)
    T = np.concatenate([[10], seq, [10]]).astype(np.uint8)
    pats = pkg.pattern_array(primers, True)
    out = {}
    for name, sp in (("index", pkg.PM_SCAN_INDEX), ("planes", pkg.PM_SCAN_PLANES)):
        sc = pkg.Scanner(k=k, indels=indels, scan_path=sp,
                         index=pkg.PM_INDEX_EXACT_HALVES if halves else _n5(pkg, dict(k=k)))
        for i, p in enumerate(pats):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_text(T)
        out[name] = sc.find_patterns()
        if name == "index":
            assert sc.stats()["index_entries"] > 100000
        sc.close()
    assert len(out["planes"]) >= 300
    assert out["index"].tobytes() == out["planes"].tobytes()


def test_chain_longer_than_the_right_guard(pkg):
    """A poly-A run of 300 kb against a poly-A primer is ONE chain of candidates.  A rank
    whose range ends inside the run must follow it past the 64 Ki right guard (the scan
    is repeated with a longer guard) to report the same hit as a whole-stream scan."""
    import casegen
    rng = random.Random(3)
    primers = ["A" * 16, casegen.rand_dna(rng, 20)]
    seq = list(casegen.rand_dna(rng, 500_000))
    seq[40_000:340_000] = "A" * 300_000
    seq[400_000:400_020] = list(primers[1])
    st = pkg.normalize_fasta((">s\n" + "".join(seq) + "\n").encode())
    T = st.chars

    def run(begin, end):
        sc = pkg.Scanner(k=1, indels=True, index=pkg.PM_INDEX_FILTER_BITVEC)
        for i, p in enumerate(primers):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_text(T)
        h = sc.find_patterns(begin, end)
        sc.close()
        return h

    whole = run(0, pkg.INT64_MAX)
    parts = [run(0, 50_000), run(50_000, 390_000), run(390_000, pkg.INT64_MAX)]
    merged = np.concatenate(parts)
    merged = merged[np.lexsort((merged["pattern_id"], merged["first"]))]
    assert len(whole) >= 2
    assert merged.tobytes() == whole.tobytes()
    assert len(parts[0]) >= 1     # the run's single hit belongs to the rank its chain starts in


def test_many_ambiguous_characters_fall_back_to_planes(pkg):
    """With -w a text 'R' matches the pattern symbols A, G and R: the pair index lists
    such characters and re-examines their windows; past 2^20 of them it hands the
    text to the bit-plane kernels.  Both routes must give the same hits."""
    import casegen
    rng = random.Random(31)
    primers = [casegen.rand_dna(rng, 20) for _ in range(6)] + ["ACGTRYACGTACGTNACGTA"]
    n = 3_000_000
    g = np.random.default_rng(5)
    seq = np.frombuffer(b"ACGTRRR", np.uint8)[g.integers(0, 7, n)]
    for j in range(60):
        q = casegen.mutate(rng, "".join(c if c in "ACGT" else "A" for c in primers[j % 7]), rng.randint(0, 1), False)
        at = 1000 + j * 40000
        seq[at:at + len(q)] = np.frombuffer(q.encode(), np.uint8)
    T = np.concatenate([[10], seq, [10]]).astype(np.uint8)
    out = {}
    for name, sp in (("auto", pkg.PM_SCAN_AUTO), ("planes", pkg.PM_SCAN_PLANES)):
        sc = pkg.Scanner(k=1, indels=False, wildcards=True, scan_path=sp, index=pkg.PM_INDEX_FILTER_BITVEC)
        for i, p in enumerate(primers):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_text(T)
        out[name] = sc.find_patterns()
        st = sc.stats()
        assert st["scan_path"] == pkg.PM_SCAN_PLANES
        if name == "auto":
            assert st["n_ambiguous"] > (1 << 20)
        sc.close()
    assert len(out["planes"]) >= 40
    assert out["auto"].tobytes() == out["planes"].tobytes()


def test_full_size_properties(pkg):
    """BASELINE-shaped run (1k 20-mers, both strands, k=2 mismatches) on 64 Mbp:
    planted copies are all found with the planted distance, hit lists are
    sorted, and the k=1 hits are a subset of the k=2 hits."""
    import casegen
    rng = random.Random(9)
    nprim = 1000
    primers = [casegen.rand_dna(rng, 20) for _ in range(nprim)]
    n = 64 << 20
    g = np.random.default_rng(3)
    seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, n)]
    planted = []
    for j in range(400):
        pi = rng.randrange(nprim)
        q = casegen.mutate(rng, primers[pi], rng.randint(0, 2), False)
        at = 1000 + j * 150000
        seq[at:at + 20] = np.frombuffer(q.encode(), np.uint8)
        planted.append((at, pi))
    T = np.concatenate([[10], seq, [10]]).astype(np.uint8)
    pats = pkg.pattern_array(primers, True)
    res = {}
    for k in (1, 2):
        sc = pkg.Scanner(k=k, indels=False)
        for i, p in enumerate(pats):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_text(T)
        h = sc.find_patterns()
        sc.close()
        assert np.all(np.diff(h["first"]) >= 0)
        res[k] = {(int(r["end"]), int(r["pattern_id"])) for r in h}
        if k == 2:
            for at, pi in planted:
                assert (at + 1 + 20, pi + 1) in res[k]
            # hits report the true number of mismatches
            for r in h[:200]:
                s, e, pid = int(r["start"]), int(r["end"]), int(r["pattern_id"])
                w = T[e - 20:e].tobytes().decode()
                assert sum(a != b for a, b in zip(w, pats[pid - 1])) == int(r["editdist"])
    assert res[1] <= res[2]


def test_device_pointer_with_4_byte_alignment(pkg):
    """pm_set_text_device only asks for 4-byte alignment: a view into a larger device buffer."""
    import torch
    import casegen
    rng = random.Random(21)
    primers = [casegen.rand_dna(rng, 20) for _ in range(20)]
    n = 300_000
    g = np.random.default_rng(4)
    seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, n)].copy()
    for j in range(50):
        seq[1000 + j * 5000:1020 + j * 5000] = np.frombuffer(primers[j % 20].encode(), np.uint8)
    T = np.concatenate([[10], seq, [10]]).astype(np.uint8)
    big = torch.zeros(len(T) + 64, dtype=torch.uint8, device="cuda")
    out = []
    for shift in (0, 4, 12):
        view = big[shift:shift + len(T)]
        view.copy_(torch.from_numpy(T))
        sc = pkg.Scanner(k=2, indels=False)
        for i, p in enumerate(primers):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_text(view)
        out.append(sc.find_patterns())
        sc.close()
    assert len(out[0]) >= 50
    assert out[1].tobytes() == out[0].tobytes() and out[2].tobytes() == out[0].tobytes()


def test_host_pipeline_equals_device_path(pkg):
    """Pinned-host text goes through the chunked copy/pack/scan pipeline (several
    chunks at this size); the result must equal the device-resident path."""
    import torch
    import casegen
    rng = random.Random(5)
    primers = [casegen.rand_dna(rng, 20) for _ in range(40)]
    n = 110 * (1 << 20) + 12345
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    d = lut[torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.uint8).long()]
    d[0] = 10
    d[-1] = 10
    for j in range(140):   # planted copies, 100 of them around the 64 MiB chunk seam
        at = (64 << 20) - 2000 + 41 * j if j < 100 else 5000 + 700001 * (j - 99)
        d[at:at + 20] = torch.tensor(list(primers[j % 40].encode()), dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.copy_(d)
    torch.cuda.synchronize()
    out = []
    for src, sp in ((d, pkg.PM_SCAN_PLANES), (h, pkg.PM_SCAN_PLANES), (h.numpy().copy(), pkg.PM_SCAN_PLANES),
                    (d, pkg.PM_SCAN_INDEX), (h, pkg.PM_SCAN_INDEX)):
        sc = pkg.Scanner(k=2, indels=False, scan_path=sp)
        for i, p in enumerate(pkg.pattern_array(primers, True)):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_text(src)
        out.append(sc.find_patterns())
        sc.close()
    assert len(out[0]) >= 140
    assert all(o.tobytes() == out[0].tobytes() for o in out[1:])


@pytest.mark.parametrize("k", [0, 2])
def test_range_end_inside_a_chunked_text(pkg, k):
    """pm_find_patterns(begin, end) on a text that arrived in several chunks (pm_set_text_host
    over 64 MiB), with `end` on / just behind tile boundaries inside the first chunk and at the
    chunk seam: the positions just before `end` must be visited although no later chunk is."""
    import torch
    import casegen
    rng = random.Random(50 + k)
    primers = [casegen.rand_dna(rng, 20) for _ in range(30)]
    n = 130 * (1 << 20) + 777
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    d = lut[torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.uint8).long()]
    d[0] = 10
    d[-1] = 10
    ends = [t * TILE_POS + e for t in (1000, 2048, 2049) for e in (0, 1, 20)]
    for t in (1000, 2048, 2049):
        base = t * TILE_POS
        for j in range(60):     # dense copies from 700 before to 500 past the boundary, odd spacing
            at = base - 700 + 21 * j
            d[at:at + 20] = torch.tensor(list(primers[j % 30].encode()), dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.copy_(d)
    torch.cuda.synchronize()
    sc = pkg.Scanner(k=k, indels=False, scan_path=pkg.PM_SCAN_INDEX)
    for i, p in enumerate(primers):
        sc.add_pattern(p, i + 1)
    sc.init()
    sc.set_text(h)
    whole = sc.find_patterns()
    assert len(whole) >= 180
    for e in ends:
        for begin in (0, 900 * TILE_POS + 5):
            sc.set_text(h)      # a fresh chunked upload: the scan follows the copy
            part = sc.find_patterns(begin, e)
            want = whole[(whole["first"] >= begin) & (whole["first"] < e)]
            assert part.tobytes() == want.tobytes(), (e, begin)
    sc.close()


@pytest.mark.parametrize("k", [2, 3])
def test_edit_scan_dense_chunks_take_the_slow_path(pkg, oracle, k):
    """Text that repeats a piece pair of some pattern over and over (poly-A against a poly-A primer,
    a dinucleotide repeat against its primer): every lane of a chunk has level-1 hits at every
    distance, the per-warp record queue of pair_scan_edit_kernel overflows and the chunk is redone by
    its slow path.  Raw candidates against the oracle's recurrence."""
    import casegen
    rng = random.Random(700 + k)
    m = 4 * (k + 2) + 5
    primers = ["A" * m, ("AC" * 20)[:m], casegen.rand_dna(rng, m), casegen.rand_dna(rng, m)]
    n = 40_000
    seq = list(casegen.rand_dna(rng, n))
    seq[3_000:9_000] = "A" * 6_000
    seq[15_000:19_000] = list(("AC" * 2_000))
    seq[19_000:19_600] = "A" * 600            # the two repeats next to each other
    for at in (25_000, 30_010, 35_555):
        q = casegen.mutate(rng, primers[2], k, True)
        seq[at:at + len(q)] = list(q)
    seq[8_990] = "N"
    st = pkg.normalize_fasta((">r\n" + "".join(seq[:20_000]) + "\n>s\n" + "".join(seq[20_000:]) + "\n").encode())
    T = st.chars
    want = sorted({(int(b), int(a)) for a, b, _ in oracle.wm_scan(T, primers, k, True, False, False)})
    sc = pkg.Scanner(k=k, indels=True, index=pkg.PM_INDEX_FILTER_BITVEC, scan_path=pkg.PM_SCAN_INDEX)
    for i, p in enumerate(primers):
        sc.add_pattern(p, i + 1)
    sc.init()
    sc.set_text(T)
    got = sorted((int(b), int(a)) for a, b in sc.scan_candidates())
    stats = sc.stats()
    hits = sc.find_patterns()
    sc.close()
    assert stats["scan_path"] == pkg.PM_SCAN_INDEX and len(want) > 9_000
    assert got == want
    # ... and the hit list against the oracle's (chains thousands of candidates long: the pool path)
    recs = pkg.hits_to_records(st, hits, len(primers))
    ref = oracle.primer_match(oracle.normalize_fasta((">r\n" + "".join(seq[:20_000]) + "\n>s\n" + "".join(seq[20_000:]) + "\n").encode()),
                              primers, k=k, indels=True, rev_comp=False)
    assert as_tuples(recs) == as_tuples(ref)


@pytest.mark.parametrize("k", [0, 1])
def test_candidate_buffer_grows_on_a_tandem_repeat(pkg, oracle, k):
    """Every position of an (ACGT)n text ends an occurrence of one of the four rotations of
    (ACGT)5: 1.2 M candidates against a buffer that starts at 2^20 entries.  The scan counts past
    the end of the buffer, the buffer grows to the count and the scan runs once more."""
    unit = "ACGT"
    primers = [(unit[r:] + unit[:r]) * 5 for r in range(4)]
    n = 1_200_000
    st = pkg.normalize_fasta((">t\n" + unit * (n // 4) + "\n").encode())
    T = st.chars
    raw = oracle.exact_scan(T, primers, False, False) if k == 0 else oracle.wm_scan(T, primers, k, False, False, False)
    want = np.array(sorted((int(b), int(a)) for a, b, _ in raw), dtype=np.int64)
    sc = pkg.Scanner(k=k, indels=False, index=_n5(pkg, dict(k=k)), scan_path=pkg.PM_SCAN_INDEX)
    for i, p in enumerate(primers):
        sc.add_pattern(p, i + 1)
    sc.init()
    sc.set_text(T)
    cands = sc.scan_candidates()
    got = np.array(sorted((int(b), int(a)) for a, b in cands), dtype=np.int64)
    hits = sc.find_patterns()
    sc.close()
    assert len(want) > (1 << 20)
    assert got.shape == want.shape and (got == want).all()
    # occurrences of a rotation lie 4 apart, further than the 2k+1 a chain bridges: one hit per candidate
    assert len(hits) == len(want)
    assert (np.diff(hits["first"].astype(np.int64)) >= 0).all()
    assert int((hits["editdist"] != 0).sum()) == 0


def test_seed_list_overflow_is_retried(pkg, oracle, monkeypatch):
    """A seed list far too small for the text (PM_SEED_LIST_ENTRIES): the scan is cut into smaller
    batches, then the list grows, and the candidates are those of the oracle's recurrence."""
    import casegen
    monkeypatch.setenv("PM_SEED_LIST_ENTRIES", "1024")
    rng = random.Random(811)
    k = 2
    m = 21
    primers = ["A" * m, casegen.rand_dna(rng, m), casegen.rand_dna(rng, m)]
    seq = list(casegen.rand_dna(rng, 90_000))
    seq[10_000:14_000] = "A" * 4_000
    seq[70_000:71_000] = "A" * 1_000
    for at in (5_000, 40_000, 80_000):
        q = casegen.mutate(rng, primers[1], k, True)
        seq[at:at + len(q)] = list(q)
    st = pkg.normalize_fasta((">r\n" + "".join(seq) + "\n").encode())
    T = st.chars
    want = sorted({(int(b), int(a)) for a, b, _ in oracle.wm_scan(T, primers, k, True, False, False)})
    for path in (pkg.PM_SCAN_INDEX, pkg.PM_SCAN_PLANES):
        sc = pkg.Scanner(k=k, indels=True, index=pkg.PM_INDEX_FILTER_BITVEC, scan_path=path)
        for i, p in enumerate(primers):
            sc.add_pattern(p, i + 1)
        sc.init()
        sc.set_text(T)
        got = sorted((int(b), int(a)) for a, b in sc.scan_candidates())
        n_seeds = sc.stats()["n_seeds"]
        sc.close()
        assert got == want, path
        # the 4 000 poly-A positions lie in one tile, the smallest batch: more seeds than the
        # list held at the start, it had to grow
        assert n_seeds > 2 * 1024, (path, n_seeds)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_edit_seeds_through_flagged_characters_are_rejected(pkg, oracle, k):
    """The pair index keys on 2-bit codes; an end-of-sequence or N character carries an arbitrary
    code and can complete a piece by accident (found by tools/fuzz_vs_reference.py: a sequence that
    begins with a primer minus its first character -- the reference reports nothing there, the
    deletion would fall into the column behind the end of sequence).  Sequences that begin or end
    with every primer minus a character, N in place of primer characters: candidates of the
    oracle's recurrence, exactly."""
    import casegen
    rng = random.Random(900 + k)
    m = 4 * (k + 2) + rng.randrange(0, 4)
    core = casegen.rand_dna(rng, m - 1)
    primers = [c + core for c in "ACGT"] + [core + c for c in "ACGT"]
    seqs = []
    for i in range(400):
        body = casegen.rand_dna(rng, rng.randrange(20, 60))
        kind = i % 4
        if kind == 0:
            s = core + body                      # first character of the primer missing at the start
        elif kind == 1:
            s = body + core                      # last character missing at the end
        elif kind == 2:
            s = body + "N" + core + body         # an N where the first character would be
        else:
            s = body + core + "N" + body
        seqs.append(s)
    fasta = "".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(seqs))
    st = pkg.normalize_fasta(fasta.encode())
    T = st.chars
    want = sorted({(int(b), int(a)) for a, b, _ in oracle.wm_scan(T, primers, k, True, False, False)})
    sc = pkg.Scanner(k=k, indels=True, index=pkg.PM_INDEX_FILTER_BITVEC, scan_path=pkg.PM_SCAN_INDEX)
    for i, p in enumerate(primers):
        sc.add_pattern(p, i + 1)
    sc.init()
    sc.set_text(T)
    got = sorted((int(b), int(a)) for a, b in sc.scan_candidates())
    sc.close()
    assert len(want) > 400
    assert got == want
