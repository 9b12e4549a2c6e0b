"""The oracle (oracle/pm_oracle.c) against the golden vectors the UNMODIFIED
reference produced (tests/golden/make_golden.py).  CPU only."""
import pytest

from conftest import as_tuples, records_of


def test_golden_has_cases(golden):
    assert len(golden["cases"]) >= 60
    assert sum(len(c["records"]) for c in golden["cases"]) > 500


def test_oracle_matches_reference_golden(oracle, golden):
    for case in golden["cases"]:
        st = oracle.normalize_fasta(case["fasta"].encode("latin-1"))
        got = oracle.primer_match(st, case["primers"], **case["params"])
        assert as_tuples(got) == sorted(records_of(case)), "case %s" % case["seed"]


def test_oracle_scan_levels_are_monotone(oracle):
    # every candidate of the k-error scan is a candidate of the (k+1)-error scan
    import casegen
    fasta, primers = casegen.make_case(5, nseq=3, seqlen=600, nprimers=8, k=2)
    st = oracle.normalize_fasta(fasta.encode())
    prev = None
    for k in (1, 2, 3):
        c = oracle.wm_scan(st.chars, primers, k, True)
        cur = {(int(a), int(b)) for a, b, _ in c}
        if prev is not None:
            assert prev <= cur
        prev = cur


def test_oracle_hamming_candidates_are_mismatch_counts(oracle):
    # substitutions only: candidate <=> at most k mismatches and no eos in the window
    import casegen
    import numpy as np
    fasta, primers = casegen.make_case(11, nseq=3, seqlen=500, nprimers=6, k=2, indels=False)
    st = oracle.normalize_fasta(fasta.encode())
    T = st.chars
    for k in (1, 2):
        got = {(int(a), int(b)) for a, b, _ in oracle.wm_scan(T, primers, k, False)}
        want = set()
        for pi, p in enumerate(primers):
            pa = np.frombuffer(p.encode(), np.uint8)
            m = len(pa)
            for t in range(m - 1, len(T)):
                w = T[t - m + 1:t + 1]
                if (w == 10).any():
                    continue
                if int((w != pa).sum()) <= k:
                    want.add((t + 1, pi + 1))
        assert got == want


def test_oracle_exact_halves_matches_reference_golden(oracle, halves_golden):
    """k = 1 with the reference's automatic index choice (exact_halves over its exact index,
    primer_alignment.cc DP), and exact_halves forced with k = 2."""
    assert len(halves_golden["cases"]) >= 40
    assert sum(len(c["records"]) for c in halves_golden["cases"]) > 300
    for case in halves_golden["cases"]:
        st = oracle.normalize_fasta(case["fasta"].encode("latin-1"))
        if case["select"] == 0:
            assert oracle.auto_index(case["primers"], case["params"]["k"]) == "exact_halves"
        got = oracle.primer_match(st, case["primers"], index="exact_halves", **case["params"])
        assert as_tuples(got) == sorted(records_of(case)), "case %s" % case["seed"]


def test_oracle_with_exact_base_constraints_matches_reference_golden(oracle, constraints_golden):
    """-s / -e / -5 / -3: the constraints inside every alignment, and the index they make
    pick_pattern_index choose (exact_bases from 6 characters on)."""
    cases = constraints_golden["cases"]
    assert len(cases) >= 40 and {c["index"] for c in cases} >= {"exact_bases", "exact_halves", "filter_bitvec"}
    for case in cases:
        st = oracle.normalize_fasta(case["fasta"].encode("latin-1"))
        got = oracle.primer_match(st, case["primers"], index="auto" if case["select"] == 0 else "filter_bitvec",
                                  **case["params"], **case["constraints"])
        assert as_tuples(got) == sorted(records_of(case)), "case %s" % case["seed"]


def test_oracle_pcr_match_matches_reference_golden(oracle, pcr_golden):
    """pcr_match (primer pairs -> amplicons) over the same scan: every record the reference
    printed, in its order; -a, -b, -m / -M, constraints, UniSTS sizes with -d."""
    cases = pcr_golden["cases"]
    assert len(cases) >= 60 and sum(1 for c in cases if "sts" in c) >= 10
    n = 0
    for case in cases:
        st = oracle.normalize_fasta(case["fasta"].encode("latin-1"))
        par = dict(case["params"])
        select = par.pop("select")
        got = oracle.pcr_match(st, case["primers"], index="auto" if select == 0 else "filter_bitvec",
                               sts_sizes=case.get("sts_sizes"), **par)
        assert got == oracle.parse_pcr_records(case["stdout_records"]), "case %s" % case["seed"]
        n += len(got)
    assert n > 300
