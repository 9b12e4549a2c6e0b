"""Generates tests/golden/large_golden.json from the UNMODIFIED reference.

    make -C oracle && python tests/golden/make_large_golden.py

A few databases of 1-3 Mbp that are too large to commit: the FASTA text is rebuilt
from a seed (large_case below, numpy PCG64), only the records the reference printed
are stored.  Several sequences, 70-column lines, planted primer copies with up to k
errors, runs of N, soft-masked (lower-case) stretches, IUPAC characters in the
text, a poly-A stretch; 40 to 200 primers, both strands; the reference's automatic
index choice (exact_halves for k = 1, filter_bitvec for k >= 2) and -N 5.
"""
from __future__ import annotations

import json
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import casegen  # noqa: E402

CASES = [
    dict(seed=1, n=1_500_000, nseq=3, nprimers=200, plen=20, k=2, indels=False, wc=False, tn=False, select=0),
    dict(seed=2, n=1_000_000, nseq=5, nprimers=60, plen=20, k=1, indels=False, wc=False, tn=False, select=0),
    dict(seed=3, n=1_000_000, nseq=2, nprimers=60, plen=24, k=1, indels=True, wc=False, tn=False, select=0),
    dict(seed=4, n=1_200_000, nseq=4, nprimers=80, plen=25, k=3, indels=True, wc=False, tn=False, select=0),
    dict(seed=5, n=1_000_000, nseq=3, nprimers=40, plen=20, k=2, indels=False, wc=True, tn=False, select=0),
    dict(seed=6, n=2_500_000, nseq=2, nprimers=100, plen=20, k=0, indels=False, wc=False, tn=False, select=0),
    dict(seed=7, n=1_000_000, nseq=3, nprimers=60, plen=20, k=1, indels=True, wc=False, tn=False, select=5),
    dict(seed=8, n=800_000, nseq=2, nprimers=30, plen=18, k=2, indels=True, wc=True, tn=True, select=0),
    # hit-dense: primer sites in tandem repeats (runs of copies 20-26 apart, so that chains of candidates
    # touch and overlap), thousands of hits per 100 kbp
    dict(seed=9, n=600_000, nseq=2, nprimers=40, plen=20, k=2, indels=False, wc=False, tn=False, select=0, dense=4000),
    dict(seed=10, n=400_000, nseq=2, nprimers=30, plen=20, k=2, indels=True, wc=False, tn=False, select=0, dense=2000),
]


def large_case(c):
    """(fasta bytes, primers) of a case: everything derives from c["seed"]."""
    g = np.random.default_rng(1000 + c["seed"])
    rng = random.Random(2000 + c["seed"])
    primers = [casegen.rand_dna(rng, c["plen"]) for _ in range(c["nprimers"])]
    if c["wc"]:
        for i in range(0, len(primers), 3):      # one to three ambiguity codes, like db/patwc.txt
            p = list(primers[i])
            for _ in range(rng.randint(1, 3)):
                p[rng.randrange(len(p))] = rng.choice("RYKMSWBDHVN")
            primers[i] = "".join(p)
    seq = np.frombuffer(b"ACGT", np.uint8)[g.integers(0, 4, c["n"])].copy()
    n = c["n"]
    for j in range(400):                          # planted copies, either strand, up to k errors
        p = primers[rng.randrange(len(primers))]
        q = "".join(ch if ch in "ACGT" else rng.choice("ACGT") for ch in p)
        if rng.random() < 0.5:
            q = q[::-1].translate(str.maketrans("ACGT", "TGCA"))
        q = casegen.mutate(rng, q, rng.randint(0, c["k"]), c["indels"])
        at = rng.randrange(100, n - 100)
        seq[at:at + len(q)] = np.frombuffer(q.encode(), np.uint8)
    if c.get("dense"):                            # tandem repeats of primer sites
        left = c["dense"]
        while left > 0:
            run = min(left, rng.randint(5, 60))
            at = rng.randrange(100, n - 100 - 30 * run)
            p = primers[rng.randrange(len(primers))]
            if rng.random() < 0.5:
                p = p[::-1].translate(str.maketrans("ACGT", "TGCA"))
            for _ in range(run):
                q = casegen.mutate(rng, p, rng.randint(0, c["k"]), c["indels"])
                seq[at:at + len(q)] = np.frombuffer(q.encode(), np.uint8)
                at += len(q) + rng.randint(0, 6)
            left -= run
    for _ in range(6):                            # runs of N
        at, ln = rng.randrange(0, n - 5000), rng.choice([1, 7, 300, 4000])
        seq[at:at + ln] = ord("N")
    for _ in range(4):                            # soft-masked stretches
        at, ln = rng.randrange(0, n - 3000), rng.choice([50, 2000])
        seq[at:at + ln] |= 0x20
    if c["wc"]:
        for _ in range(60):                       # IUPAC characters in the text
            seq[rng.randrange(n)] = ord(rng.choice("RYKMSWBDHV"))
    at = rng.randrange(0, n - 200)
    seq[at:at + 120] = ord("A")                   # a poly-A stretch
    cuts = sorted(rng.sample(range(1000, n - 1000), c["nseq"] - 1)) if c["nseq"] > 1 else []
    bounds = [0] + cuts + [n]
    parts = []
    for i in range(len(bounds) - 1):
        body = seq[bounds[i]:bounds[i + 1]]
        full = len(body) // 70 * 70
        lines = np.concatenate([body[:full].reshape(-1, 70), np.full((full // 70, 1), 10, np.uint8)], axis=1)
        parts += [b">large%d_%d synthetic\n" % (c["seed"], i), lines.tobytes(), body[full:].tobytes(), b"\n"]
    return b"".join(parts), primers


def params_of(c):
    return dict(k=c["k"], indels=c["indels"], wc=c["wc"], tn=c["tn"], rev_comp=True)


def main():
    import pm_oracle as O
    if not O.ref_available():
        raise SystemExit("oracle/_ref/primer_match missing: run `make -C oracle` first")
    tmp = tempfile.mkdtemp()
    out_cases = []
    out = os.path.join(HERE, "large_golden.json")
    have = {}
    if os.path.exists(out) and "--all" not in sys.argv:      # keep what is there, add the new cases
        with open(out) as f:
            have = {e["case"]["seed"]: e for e in json.load(f)["cases"]}
    for c in CASES:
        if c["seed"] in have and have[c["seed"]]["case"] == c:
            out_cases.append(have[c["seed"]])
            continue
        fasta, primers = large_case(c)
        path = os.path.join(tmp, "large%d.fasta" % c["seed"])
        with open(path, "wb") as f:
            f.write(fasta)
        recs = O.run_reference(path, primers, select=c["select"], timeout=3600, **params_of(c))
        os.unlink(path)
        out_cases.append(dict(case=c, records=[[r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index,
                                                r.editdist, r.stream_start, r.stream_end, r.alignment] for r in recs]))
        print("case %d: %d records" % (c["seed"], len(recs)), flush=True)
    with open(out, "w") as f:
        json.dump(dict(generator="tests/golden/make_large_golden.py", cases=out_cases), f, indent=0)
    print("wrote", out)


if __name__ == "__main__":
    main()
