"""k = 4 and 5 (-k / -K): error counts beyond the pair index's k <= 3, served by the bit-plane
kernels.  The records the UNMODIFIED reference printed (tests/golden/k45_golden.json) pin the oracle
(not gpu) and the CUDA path (gpu); seeded cases compare the CUDA path with the oracle."""
import json
import os
import random

import pytest

from conftest import ROOT, as_tuples


@pytest.fixture(scope="module")
def k45_golden():
    with open(os.path.join(ROOT, "tests", "golden", "k45_golden.json")) as f:
        return json.load(f)


def test_oracle_matches_reference_for_k4_k5(oracle, k45_golden):
    n = 0
    for case in k45_golden["cases"]:
        got = oracle.primer_match(oracle.normalize_fasta(case["fasta"].encode()), case["primers"], **case["params"])
        assert as_tuples(got) == sorted(tuple(r) for r in case["records"]), "case %d" % case["seed"]
        n += len(got)
    assert n >= 40


@pytest.mark.gpu
def test_device_matches_reference_for_k4_k5(pkg, built_lib, k45_golden):
    for case in k45_golden["cases"]:
        st = pkg.normalize_fasta(case["fasta"].encode())
        got = pkg.primer_match(st, case["primers"], index=pkg.PM_INDEX_FILTER_BITVEC, **case["params"])
        assert as_tuples(got) == sorted(tuple(r) for r in case["records"]), "case %d" % case["seed"]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6200, 6216))
def test_k4_k5_match_oracle_seeded(pkg, built_lib, oracle, seed):
    import casegen
    rng = random.Random(seed)
    k = rng.choice([4, 5])
    indels = rng.random() < 0.5
    wc = rng.random() < 0.3
    params = dict(k=k, indels=indels, wc=wc, tn=wc and rng.random() < 0.5, rev_comp=rng.random() < 0.6)
    kw = dict(nseq=rng.choice([1, 3]), seqlen=rng.choice([400, 1500]), nprimers=rng.choice([2, 6]),
              plen=rng.choice([(12, 20), (20, 32), (33, 64)]), k=k, indels=indels, wc=wc,
              text_iupac=rng.choice([0, 0.01]), lower=0, repeats=rng.random() < 0.5, line=60, crlf=False)
    fasta, primers = casegen.make_case(seed, **kw)
    raw = fasta.encode()
    want = oracle.primer_match(oracle.normalize_fasta(raw), primers, **params)
    got = pkg.primer_match(pkg.normalize_fasta(raw), primers, index=pkg.PM_INDEX_FILTER_BITVEC, **params)
    assert as_tuples(got) == as_tuples(want)
