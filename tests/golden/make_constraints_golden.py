"""Generates tests/golden/constraints_golden.json from the UNMODIFIED reference.

    make -C oracle && python tests/golden/make_constraints_golden.py

primer_match's -s / -e / -5 / -3 options force the first / last characters of a primer to
match exactly (`~n`: all but n).  They reach every alignment as exact_start_bases /
exact_end_bases (primer_match.cc:990-1060) and they steer pick_pattern_index: constraints
of 6 or more characters select exact_bases (select.cc:131-133, exact_bases.cc), short
ones leave exact_halves (k = 1) or filter_bitvec.  Half of the cases use the reference's
automatic choice, half -N 5.  Seeds on which the reference aborts are skipped.
"""
from __future__ import annotations

import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import casegen  # noqa: E402
import pm_oracle as O  # noqa: E402


def cli_constraints(cons):
    extra = []
    for opt, key in (("-s", "stlen"), ("-e", "edlen"), ("-5", "fplen"), ("-3", "tplen")):
        v = cons[key]
        if v:
            extra += [opt, ("~%d" % -v) if v < 0 else str(v)]
    return extra


def main():
    if not O.ref_available():
        raise SystemExit("oracle/_ref/primer_match missing: run `make -C oracle` first")
    cases, skipped, seed = [], 0, -1
    tmp = tempfile.mkdtemp()
    while len(cases) < 48:
        seed += 1
        rng = random.Random(9300 + seed)
        k = rng.choice([1, 1, 2, 3])
        indels = rng.random() < 0.5
        wc = rng.random() < 0.25
        tn = wc and rng.random() < 0.5
        params = dict(k=k, indels=indels, wc=wc, tn=tn, rev_comp=rng.random() < 0.7)
        kw = dict(nseq=rng.choice([1, 3]), seqlen=rng.choice([400, 1500]), nprimers=rng.choice([2, 6, 12]),
                  plen=rng.choice([(14, 22), (20, 20), (25, 32)]), k=k, indels=indels, wc=wc,
                  text_iupac=rng.choice([0, 0, 0.02]), lower=0, repeats=rng.random() < 0.4, line=60, crlf=False)
        fasta, primers = casegen.make_case(9300 + seed, **kw)
        cons = dict(stlen=0, edlen=0, fplen=0, tplen=0)
        which = rng.choice(["fplen", "tplen", "stlen", "edlen", "both"])
        if which == "both":
            cons["fplen"], cons["tplen"] = rng.choice([3, 6, 8]), rng.choice([3, 6, 9])
        else:
            cons[which] = rng.choice([3, 6, 8, 10, -4, -8])
        select = 0 if len(cases) % 2 == 0 else 5
        path = os.path.join(tmp, "k%d.fasta" % seed)
        with open(path, "w", newline="") as f:
            f.write(fasta)
        try:
            recs = O.run_reference(path, primers, select=select, extra=cli_constraints(cons), **params)
        except RuntimeError:
            skipped += 1
            continue
        pats = O.pattern_array(primers, params["rev_comp"])
        lc, rc = O.pattern_constraints(primers, params["rev_comp"], **cons)
        cases.append(dict(seed=seed, params=params, constraints=cons, select=select,
                          index=O.auto_index(pats, k, lc, rc) if select == 0 else "filter_bitvec",
                          fasta=fasta, primers=primers,
                          records=[[r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index, r.editdist,
                                    r.stream_start, r.stream_end, r.alignment] for r in recs]))
    out = os.path.join(HERE, "constraints_golden.json")
    with open(out, "w") as f:
        json.dump(dict(generator="tests/golden/make_constraints_golden.py", skipped_reference_aborts=skipped,
                       cases=cases), f, indent=0)
    from collections import Counter
    print("wrote %s: %d cases, %d records, %d reference aborts skipped, indexes %s" % (
        out, len(cases), sum(len(c["records"]) for c in cases), skipped, dict(Counter(c["index"] for c in cases))))


if __name__ == "__main__":
    main()
