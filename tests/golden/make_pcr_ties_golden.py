"""Generates tests/golden/pcr_ties_golden.json from the UNMODIFIED reference pcr_match.

    make -C oracle && python tests/golden/make_pcr_ties_golden.py

Cases of the differential campaign (tools/fuzz_vs_reference.py --mode pcr) in which hits of
different primers end at the same position and the list is longer than 16: the reference's two
std::sort calls on the position alone (filter_bitvec.cc:91 on the candidates, pcr_match.cc:958
on the hits) leave such hits in an order of their own, and which of two partners comes first
decides whether they pair.  Every one of these seeds printed different records under "equal
positions by pattern id" or under "equal positions in the order reported".
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

import make_pcr_golden  # noqa: E402
import pm_oracle as O  # noqa: E402

SEEDS = [40082, 40536, 41789, 42589, 43517, 43738, 44671, 44673, 44689, 44894, 44926, 45145, 45714, 45900, 45922]


def fuzz_case(seed: int):
    """The pcr case of tools/fuzz_vs_reference.py for this seed."""
    rng = random.Random(seed * 13 + 5)
    k = rng.choice([0, 1, 1, 2, 2, 3])
    indels = rng.random() < 0.5
    wc = rng.random() < 0.2
    rev_comp = rng.random() < 0.7
    params = dict(k=k, indels=indels, wc=wc, tn=wc and rng.random() < 0.5, rev_comp=rev_comp,
                  select=rng.choice([0, 0, 5]) if k else 0,
                  mindist=rng.choice([0, 0, 0, 60, 200]), maxdist=rng.choice([2000, 2000, 700, 250]),
                  betweenlen=rng.random() < 0.25, allorient=rng.random() < 0.35,
                  stlen=0, edlen=0, fplen=0, tplen=0)
    if k and rng.random() < 0.3:
        params[rng.choice(["stlen", "edlen", "fplen", "tplen"])] = rng.choice([3, 6, 8, -5])
    fasta, primers = make_pcr_golden.make_pcr_case(seed, k, indels, wc, rev_comp)
    return params, fasta, primers


def main():
    if not os.path.exists(O.PCR_REF_BIN):
        raise SystemExit("oracle/_ref/pcr_match missing: run `make -C oracle` first")
    tmp = tempfile.mkdtemp()
    cases = []
    for seed in SEEDS:
        params, fasta, primers = fuzz_case(seed)
        path = os.path.join(tmp, "t%d.fasta" % seed)
        with open(path, "w", newline="") as f:
            f.write(fasta)
        cases.append(dict(seed=seed, params=params, fasta=fasta, primers=primers,
                          stdout_records=O.run_reference_pcr(path, primers, **params)))
    out = os.path.join(HERE, "pcr_ties_golden.json")
    with open(out, "w") as f:
        json.dump(dict(generator="tests/golden/make_pcr_ties_golden.py", cases=cases), f, indent=0)
    print("wrote %s: %d cases, %d records" % (out, len(cases), sum(len(c["stdout_records"].splitlines()) for c in cases)))


if __name__ == "__main__":
    main()
