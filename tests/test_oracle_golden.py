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


def test_oracle_pcr_match_reproduces_the_order_of_simultaneous_hits(oracle, pcr_ties_golden):
    """Hits of different primers that end at one position: the reference leaves them in the order
    its two std::sort calls on the position alone leave them (filter_bitvec.cc:91, pcr_match.cc:958),
    and a hit pairs only with partners that come later.  The oracle restates std::sort
    (std_sort_by_key); 15 cases of the differential campaign in which any simpler rule prints
    other records."""
    import dataclasses
    n = 0
    for case in pcr_ties_golden["cases"]:
        st = oracle.normalize_fasta(case["fasta"].encode("latin-1"))
        par = dict(case["params"])
        select = par.pop("select")
        got = oracle.pcr_match(st, case["primers"], index="auto" if select == 0 else "filter_bitvec", **par)
        want = oracle.parse_pcr_records(case["stdout_records"])
        assert [dataclasses.astuple(r) for r in got] == [dataclasses.astuple(r) for r in want], "case %s" % case["seed"]
        n += len(want)
    assert n > 400


def test_std_sort_restatement_matches_libstdcxx(oracle, tmp_path):
    """std_sort_by_key against the real std::sort (a 10-line C++ program compiled here) on arrays
    full of equal keys, 1 to 1000 long."""
    import random
    import subprocess
    src = tmp_path / "s.cc"
    src.write_text("#include <algorithm>\n#include <cstdio>\n#include <vector>\n"
                   "struct E { long k; long v; bool operator<(const E &o) const { return k < o.k; } };\n"
                   "int main() { long n; while (scanf(\"%ld\", &n) == 1) { std::vector<E> a(n);\n"
                   "  for (long i = 0; i < n; i++) { if (scanf(\"%ld\", &a[i].k) != 1) return 1; a[i].v = i; }\n"
                   "  std::sort(&a[0], &a[0] + n); for (auto &e : a) printf(\"%ld \", e.v); printf(\"\\n\"); } }\n")
    exe = tmp_path / "s"
    subprocess.check_call(["g++", "-O2", "-o", str(exe), str(src)])
    rng = random.Random(5)
    cases, lines = [], []
    for t in range(600):
        n = rng.choice([1, 2, 5, 16, 17, 18, 30, 33, 61, 100, 257, 1000])
        ks = [rng.randrange(max(2, n // rng.choice([1, 2, 3, 10]))) for _ in range(n)]
        if t % 7 == 0:
            ks = sorted(ks)
        if t % 11 == 0:
            ks = [0] * n
        cases.append(ks)
        lines.append("%d %s" % (n, " ".join(map(str, ks))))
    out = subprocess.run([str(exe)], input="\n".join(lines).encode(), capture_output=True, timeout=120).stdout.decode()
    for ks, line in zip(cases, out.strip().split("\n")):
        want = [int(x) for x in line.split()]
        got = [i for i, _ in oracle.std_sort_by_key(list(enumerate(ks)), lambda e: e[1])]
        assert got == want
