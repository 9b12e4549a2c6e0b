"""compress_seq's compressed database format (.tbz + .sqz) on both sides of the
scan: the oracle's restatement and the host-side encoder against the bytes the
UNMODIFIED reference compress_seq wrote (tests/golden/sqz_golden.json), and the
device path (pm_set_text_sqz_host -> unpack_sqz_kernel) against the plain-stream
path and the reference's records."""
import base64
import json
import os
import random

import numpy as np
import pytest

from conftest import ROOT, as_tuples, records_of


@pytest.fixture(scope="module")
def sqz_golden():
    with open(os.path.join(ROOT, "tests", "golden", "sqz_golden.json")) as f:
        g = json.load(f)
    for c in g["cases"]:
        c["tbz_bytes"] = base64.b64decode(c["tbz"])
        c["sqz_bytes"] = base64.b64decode(c["sqz"])
        c["tbl_bytes"] = base64.b64decode(c["tbl"])
        c["sqn_bytes"] = base64.b64decode(c["sqn"])
    return g


def test_oracle_sqz_matches_reference_files(oracle, sqz_golden):
    assert len(sqz_golden["cases"]) >= 10
    for c in sqz_golden["cases"]:
        st = oracle.normalize_fasta(c["fasta"].encode("latin-1"), uc=True)
        assert len(st.chars) == c["n_chars"]
        assert oracle.sqz_table(st.chars) == c["tbz_bytes"]
        assert oracle.sqz_encode(st.chars, c["tbz_bytes"]) == c["sqz_bytes"]
        # the normalised form: one table index per character (Normalized<>, char_io.t:216-283)
        assert c["tbl_bytes"] == c["tbz_bytes"] and len(c["sqn_bytes"]) == c["n_chars"]
        assert bytes(c["tbl_bytes"][b] for b in c["sqn_bytes"]) == st.chars.tobytes()
        dec = oracle.sqz_decode(c["sqz_bytes"], c["tbz_bytes"])
        assert dec[:c["n_chars"]].tobytes() == st.chars.tobytes()
        assert (dec[c["n_chars"]:] == 10).all()          # the writer pads with eos characters


def test_host_encoder_matches_reference_files(pkg, sqz_golden):
    for c in sqz_golden["cases"]:
        st = pkg.normalize_fasta(c["fasta"].encode("latin-1"), uc=True)
        assert pkg.sqz_table(st.chars) == c["tbz_bytes"]
        assert pkg.sqz_encode(st.chars, c["tbz_bytes"]).tobytes() == c["sqz_bytes"]


@pytest.mark.parametrize("table_size", [2, 3, 5, 6, 16, 17, 40, 129, 255])
def test_encoders_agree_for_every_width(pkg, oracle, table_size):
    rng = np.random.default_rng(table_size)
    table = bytes([10] + [b for b in rng.permutation(np.arange(1, 256))[:table_size - 1] if b != 10][:table_size - 1])
    table = table + bytes(b for b in range(256) if b not in table)[:table_size - len(table)]
    chars = np.frombuffer(table, np.uint8)[rng.integers(0, len(table), 1000)]
    enc = pkg.sqz_encode(chars, table)
    assert enc.tobytes() == oracle.sqz_encode(chars, table)
    assert oracle.sqz_decode(enc.tobytes(), table, len(chars)).tobytes() == chars.tobytes()


# --------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_sqz_records_match_reference(pkg, built_lib, sqz_golden):
    for c in sqz_golden["cases"]:
        st = pkg.normalize_fasta(c["fasta"].encode("latin-1"), uc=True)
        sqz = np.frombuffer(c["sqz_bytes"], np.uint8)
        # the goldens were printed with -N 5 for k > 0 (pm_oracle.run_reference)
        n5 = pkg.PM_INDEX_FILTER_BITVEC if c["params"]["k"] > 0 else pkg.PM_INDEX_AUTO
        got = pkg.primer_match(st, c["primers"], sqz=(sqz, c["tbz_bytes"], c["n_chars"]), index=n5, **c["params"])
        assert as_tuples(got) == sorted(records_of(c)), "case %s" % c["seed"]
        got0 = pkg.primer_match(st, c["primers"], sqz=(sqz, c["tbz_bytes"], 0), index=n5, **c["params"])   # padding scanned too
        assert as_tuples(got0) == sorted(records_of(c))
        sqn = np.frombuffer(c["sqn_bytes"], np.uint8)                                  # .sqn / .tbl
        got1 = pkg.primer_match(st, c["primers"], sqz=(sqn, c["tbl_bytes"]), index=n5, **c["params"])
        assert as_tuples(got1) == sorted(records_of(c))


@pytest.mark.gpu
@pytest.mark.parametrize("table_size", [2, 4, 5, 9, 30, 100, 255])
def test_sqz_unpack_every_width(pkg, built_lib, table_size):
    """Hits through the compressed path equal hits through the plain path for every
    code width 1..8 bits (the text alphabet is padded with junk characters)."""
    import casegen
    rng = random.Random(table_size)
    primers = [casegen.rand_dna(rng, 18) for _ in range(5)]
    alphabet = (b"\nACGT" + bytes(b for b in range(33, 127) if b not in b"ACGT") +
                bytes(range(128, 256)) + bytes(range(0, 10)) + bytes(range(11, 33)) + b"\x7f")[:table_size]
    g = np.random.default_rng(table_size)
    n = 70001
    seq = np.frombuffer(alphabet, np.uint8)[g.integers(1 if table_size > 1 else 0, len(alphabet), n)].copy()
    if table_size >= 5:
        seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, n)].copy()
        junk = g.integers(0, n, 300)
        seq[junk] = np.frombuffer(alphabet, np.uint8)[g.integers(0, len(alphabet), 300)]
        for j in range(50):
            q = casegen.mutate(rng, rng.choice(primers), rng.randint(0, 1), False)
            at = 100 + j * 1300
            seq[at:at + len(q)] = np.frombuffer(q.encode(), np.uint8)
    T = np.concatenate([[10], seq, [10]]).astype(np.uint8)
    table = pkg.sqz_table(T, dna_opt=(table_size % 2 == 1))
    enc = pkg.sqz_encode(T, table)
    res = []
    for mode in ("plain", "sqz"):
        sc = pkg.Scanner(k=1, indels=False)
        for i, p in enumerate(primers):
            sc.add_pattern(p, i + 1)
        sc.init()
        if mode == "plain":
            sc.set_text(T)
        else:
            sc.set_text_sqz(enc, table, len(T))
        res.append(sc.find_patterns())
        sc.close()
    if table_size >= 5:
        assert len(res[0]) >= 50
    assert res[0].tobytes() == res[1].tobytes()


@pytest.mark.gpu
def test_sqz_pipeline_several_chunks(pkg, built_lib):
    """A compressed source of several 24 MiB chunks: unpack and pack follow the
    transfer chunk by chunk; the result equals the device-resident plain path."""
    import torch
    import casegen
    rng = random.Random(12)
    primers = [casegen.rand_dna(rng, 20) for _ in range(30)]
    n = 200_000_000 + 77
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    d = lut[torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.uint8).long()]
    d[0] = 10
    d[-1] = 10
    d[50_000_000] = 10
    seam = (24 << 20) * 8 // 3          # the character where the first source chunk ends
    for j in range(120):
        at = seam - 1500 + 29 * j if j < 100 else 9000 + 9_000_001 * (j - 99)
        d[at:at + 20] = torch.tensor(list(primers[j % 30].encode()), dtype=torch.uint8, device="cuda")
    table = b"ACGT\n"
    enc = pkg.sqz_encode_torch(d, table)
    h = torch.empty(enc.numel(), dtype=torch.uint8, pin_memory=True)
    h.copy_(enc)
    torch.cuda.synchronize()
    out = []
    for mode in ("plain", "sqz_pinned", "sqz_pageable"):
        sc = pkg.Scanner(k=2, indels=False)
        for i, p in enumerate(pkg.pattern_array(primers, True)):
            sc.add_pattern(p, i + 1)
        sc.init()
        if mode == "plain":
            sc.set_text(d)
        elif mode == "sqz_pinned":
            sc.set_text_sqz(h, table, n)
        else:
            sc.set_text_sqz(h.numpy().copy(), table, n)
        out.append(sc.find_patterns())
        sc.close()
    assert len(out[0]) >= 120
    assert out[1].tobytes() == out[0].tobytes()
    assert out[2].tobytes() == out[0].tobytes()
