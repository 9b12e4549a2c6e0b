"""Databases of 1-3 Mbp against the records the UNMODIFIED reference printed for them
(tests/golden/make_large_golden.py rebuilds the FASTA text from a seed; only the
records are committed): several sequences, N runs, soft-masked stretches, IUPAC
text, poly-A, 30-200 primers on both strands, the reference's automatic index choice."""
import json
import os

import pytest

from conftest import ROOT, as_tuples

import make_large_golden as M


@pytest.fixture(scope="module")
def large_golden():
    with open(os.path.join(ROOT, "tests", "golden", "large_golden.json")) as f:
        return json.load(f)


def _want(entry):
    return sorted(tuple(r) for r in entry["records"])


def test_oracle_matches_reference_on_large_databases(oracle, large_golden):
    assert len(large_golden["cases"]) >= 8
    for entry in large_golden["cases"]:
        c = entry["case"]
        fasta, primers = M.large_case(c)
        idx = "exact_halves" if (c["select"] == 0 and c["k"] == 1) else "filter_bitvec"
        assert idx == ("filter_bitvec" if c["select"] == 5 else oracle.auto_index(primers, c["k"])) or c["k"] == 0
        got = oracle.primer_match(oracle.normalize_fasta(fasta), primers, index=idx, **M.params_of(c))
        assert as_tuples(got) == _want(entry), "case %d" % c["seed"]


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["fasta_on_device", "stream"])
def test_device_matches_reference_on_large_databases(pkg, built_lib, large_golden, source):
    for entry in large_golden["cases"]:
        c = entry["case"]
        fasta, primers = M.large_case(c)
        p = M.params_of(c)
        pats = pkg.pattern_array(primers, True)
        sc = pkg.Scanner(k=p["k"], indels=p["indels"], wildcards=p["wc"], text_n=p["tn"], index=c["select"])
        for i, q in enumerate(pats):
            sc.add_pattern(q, i + 1)
        sc.init()
        if source == "fasta_on_device":
            st = sc.set_text_fasta(fasta)
        else:
            st = pkg.normalize_fasta(fasta)
            sc.set_text(st.chars)
        hits = sc.find_patterns()
        if c["select"] == 0 and c["k"] > 0:   # the automatic choice went where pick_pattern_index goes
            assert (sc.stats()["index_pieces"] == 2) == (c["k"] == 1) or sc.stats()["scan_path"] == pkg.PM_SCAN_PLANES
        sc.close()
        got = pkg.hits_to_records(st, hits, len(primers))
        assert as_tuples(got) == _want(entry), "case %d (%s)" % (c["seed"], source)
