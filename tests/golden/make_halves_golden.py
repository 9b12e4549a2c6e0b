"""Generates tests/golden/exact_halves_golden.json from the UNMODIFIED reference.

    make -C oracle && python tests/golden/make_halves_golden.py

For one mismatch / edit primer_match does not pick filter_bitvec: pick_pattern_index
(select.cc:120-126) wraps its exact index in exact_halves (exact_halves.cc), which
extends exact occurrences of the pattern halves with primer_alignment (the DP of
primer_alignment.cc).  Its hit list differs from filter_bitvec's on repeats and on
some indel placements, so it gets goldens of its own: the records the reference prints
with its AUTOMATIC index choice (no -N) for k = 1, plus a few runs that force
exact_halves (-N 13, or -N 14 under wildcards) with k = 2.  Seeds on which the
reference aborts (make_golden.py) are skipped.
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


def main():
    if not O.ref_available():
        raise SystemExit("oracle/_ref/primer_match missing: run `make -C oracle` first")
    cases, skipped, seed = [], 0, -1
    tmp = tempfile.mkdtemp()
    while len(cases) < 48:
        seed += 1
        rng = random.Random(7100 + seed)
        forced_k2 = len(cases) >= 40
        k = 2 if forced_k2 else 1
        indels = rng.random() < 0.5
        wc = rng.random() < 0.35
        tn = wc and rng.random() < 0.5
        params = dict(k=k, indels=indels, wc=wc, tn=tn, rev_comp=rng.random() < 0.7)
        kw = dict(nseq=rng.choice([1, 3, 5]), seqlen=rng.choice([60, 400, 1500]), nprimers=rng.choice([1, 2, 6, 12]),
                  plen=rng.choice([(8, 12), (12, 22), (20, 20), (25, 32)]), k=k, indels=indels, wc=wc,
                  text_iupac=rng.choice([0, 0, 0.01, 0.05]), lower=0, repeats=rng.random() < 0.6,
                  line=rng.choice([60, 70, 13]), crlf=rng.random() < 0.15)
        fasta, primers = casegen.make_case(7100 + seed, **kw)
        path = os.path.join(tmp, "h%d.fasta" % seed)
        with open(path, "w", newline="") as f:
            f.write(fasta)
        select = (14 if wc else 13) if forced_k2 else 0      # 0: no -N, the automatic choice
        try:
            recs = O.run_reference(path, primers, select=select, **params)
        except RuntimeError:
            skipped += 1
            continue
        cases.append(dict(seed=seed, params=params, select=select, fasta=fasta, primers=primers,
                          records=[[r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index, r.editdist,
                                    r.stream_start, r.stream_end, r.alignment] for r in recs]))
    out = os.path.join(HERE, "exact_halves_golden.json")
    with open(out, "w") as f:
        json.dump(dict(generator="tests/golden/make_halves_golden.py", skipped_reference_aborts=skipped,
                       cases=cases), f, indent=0)
    print("wrote %s: %d cases, %d records, %d reference aborts skipped" % (
        out, len(cases), sum(len(c["records"]) for c in cases), skipped))


if __name__ == "__main__":
    main()
