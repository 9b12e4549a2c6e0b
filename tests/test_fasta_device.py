"""FASTA normalisation on the device (pm_set_text_fasta_host, pm_fasta.cuh) against the
oracle's restatement of StreamedFastaFile::getch (fasta_io.t:478-629) and against
the host-side pm_normalize_fasta: stream bytes, sequence keys and header ranges
must be identical, and so must the hit records of a scan that starts from raw
FASTA bytes."""
import random

import numpy as np
import pytest

from conftest import as_tuples, records_of

pytestmark = pytest.mark.gpu


def _device_stream(pkg, raw, uc=False):
    sc = pkg.Scanner(k=0)
    sc.add_pattern("ACGTACGTAC", 1)
    sc.init()
    st = sc.set_text_fasta(raw, uc=uc, fetch_chars=True)
    sc.close()
    return st


def _check(pkg, oracle, raw, uc=False):
    want = oracle.normalize_fasta(raw, uc=uc)
    host = pkg.normalize_fasta(raw, uc=uc)
    got = _device_stream(pkg, raw, uc)
    assert host.chars.tobytes() == want.chars.tobytes()
    assert got.chars.tobytes() == want.chars.tobytes()
    assert got.seq_key.tolist() == want.seq_key.tolist()
    assert got.headers == want.headers


EDGE_FILES = [
    b"", b"\n", b"\n\n  \n", b">", b">x", b">x\n", b">x\nA", b">x\nACGT", b">x\nACGT\n", b"ACGT", b"ACGT\n>y\nTT\n",
    b">a\n>b\nACGT\n", b">a\n\n>b\nACGT\n>c\n", b">a\r\nAC\r\nGT\r\n>b\r\nTT\r\n", b">a\rACGT\r>b\rTT",
    b">a\r>b\nAC", b"\n\n>a desc more\nAC GT\n\nAC\n>b\n\n\n>c\nT\n\n", b">a\nAC>mid header\nGT\n", b">a\nacgtN\n",
    b">only header no newline", b">h1\n" + b"A" * 300 + b"\n>h2\n" + b"C" * 257 + b"\n", b">h\n" + b"\n" * 700 + b"ACGT",
    b"   >x\nAC\n", b">x\n>\n>\nA\n>\n",
]


def test_edge_files(pkg, oracle, built_lib):
    for raw in EDGE_FILES:
        for uc in (False, True):
            _check(pkg, oracle, raw, uc)


@pytest.mark.parametrize("seed", range(12))
def test_random_files(pkg, oracle, built_lib, seed):
    """Random mixtures of headers, line lengths, line terminators, blanks and '>' inside
    lines; segment (256 B) and block (32 KB) seams are crossed many times."""
    rng = random.Random(seed)
    parts = []
    for s in range(rng.choice([1, 3, 40])):
        term = rng.choice(["\n", "\n", "\r\n", "\r"])
        hdr = ">" + "".join(rng.choice("abc |>12") for _ in range(rng.choice([0, 3, 40, 300])))
        parts.append(hdr + term)
        line = rng.choice([60, 70, 13, 255, 256, 257, 1000])
        body = "".join(rng.choice("ACGTNacgt") for _ in range(rng.choice([0, 1, 500, 40000])))
        for i in range(0, len(body), line):
            parts.append(body[i:i + line] + rng.choice([term, term, term + term, " " + term]))
        if rng.random() < 0.2:
            parts.append(rng.choice([" ", "\n\n", ">"]))
    raw = "".join(parts).encode()
    if rng.random() < 0.3:
        raw = raw.rstrip()
    _check(pkg, oracle, raw, uc=bool(seed & 1))


def test_golden_records_from_raw_fasta(pkg, built_lib, golden):
    """The reference's records, starting from the raw FASTA bytes of the case."""
    for case in golden["cases"][:40]:
        raw = case["fasta"].encode("latin-1")
        p = case["params"]
        pats = pkg.pattern_array(case["primers"], p["rev_comp"])
        sc = pkg.Scanner(k=p["k"], indels=p["indels"], wildcards=p["wc"], text_n=p["tn"],
                         index=pkg.PM_INDEX_FILTER_BITVEC if p["k"] > 0 else pkg.PM_INDEX_AUTO)   # goldens: -N 5
        for i, q in enumerate(pats):
            sc.add_pattern(q, i + 1)
        sc.init()
        st = sc.set_text_fasta(raw)
        hits = sc.find_patterns()
        sc.close()
        got = pkg.hits_to_records(st, hits, len(case["primers"]))
        assert as_tuples(got) == sorted(records_of(case)), "case %s" % case["seed"]


def test_large_file_pinned_and_pageable(pkg, built_lib):
    import torch
    g = np.random.default_rng(1)
    nseq, per = 40, 1_000_000
    body = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, nseq * per)].reshape(nseq, per // 80, 80)
    lines = np.concatenate([body, np.full((nseq, per // 80, 1), 10, np.uint8)], axis=2).reshape(nseq, -1)
    raw = b"".join(b">seq%d some description\n" % i + lines[i].tobytes() for i in range(nseq))
    want = pkg.normalize_fasta(raw)
    hp = torch.empty(len(raw), dtype=torch.uint8, pin_memory=True)
    hp.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    for src in (raw, hp):
        got = _device_stream(pkg, src)
        assert got.chars.tobytes() == want.chars.tobytes()
        assert got.seq_key.tolist() == want.seq_key.tolist()
        assert got.headers == want.headers


@pytest.mark.gpu
def test_database_reader_equals_memory_path(pkg, built_lib, tmp_path):
    """pm_set_text_database: a FASTA file of several reader chunks (> 64 MiB) read through the ring of
    page-locked buffers gives the stream, sequence table, headers and hits of the in-memory path."""
    import random
    import numpy as np
    import casegen
    rng = random.Random(12)
    primers = [casegen.rand_dna(rng, 20) for _ in range(50)]
    g = np.random.default_rng(6)
    parts = []
    total = 0
    i = 0
    while total < 150 * (1 << 20):
        n = int(g.integers(2_000_000, 30_000_000))
        seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, n)].copy()
        for j in range(20):
            at = int(g.integers(0, n - 30))
            seq[at:at + 20] = np.frombuffer(primers[(i + j) % 50].encode(), np.uint8)
        body = seq[:n // 70 * 70].reshape(-1, 70)
        lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).reshape(-1)
        parts += [np.frombuffer(b">chr%d synthetic %d\n" % (i, n), np.uint8), lines]
        total += n
        i += 1
    raw = np.concatenate(parts)
    path = tmp_path / "big.fasta"
    raw.tofile(str(path))
    out = {}
    for how in ("memory", "reader"):
        sc = pkg.Scanner(k=1, indels=False, index=pkg.PM_INDEX_FILTER_BITVEC)
        for j, p in enumerate(pkg.pattern_array(primers, True)):
            sc.add_pattern(p, j + 1)
        sc.init()
        st = sc.set_text_fasta(raw) if how == "memory" else sc.set_text_database(str(path), 1)
        hits = sc.find_patterns()
        n_stream = sc.stats()["text_len"]
        head = sc.get_text(0, 100000).tobytes()
        tail = sc.get_text(n_stream - 100000, 100000).tobytes()
        out[how] = (st.seq_key.tolist(), st.headers, hits.tobytes(), n_stream, head, tail)
        sc.close()
    assert len(out["memory"][0]) == i and len(out["memory"][2]) > 100 * 120
    assert out["reader"] == out["memory"]
