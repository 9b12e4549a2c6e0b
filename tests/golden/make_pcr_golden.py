"""Generates tests/golden/pcr_match_golden.json from the UNMODIFIED reference pcr_match.

    make -C oracle && python tests/golden/make_pcr_golden.py

pcr_match (pcr_match.cc) is the second caller of the PatternMatch scan: primers come in
pairs, both strands are always searched, and hits of the two primers of a pair that lie
within -m / -M of each other on one sequence are reported as an amplicon.  Cases: planted
amplicons on both strands and in all four primer orientations (-a), exact and mutated
primer sites, spacers from 0 to past -M, amplicons that would span two sequences, N runs
inside amplicons, -b, -m, -M, -s/-e/-5/-3, IUPAC primers, primers from a UniSTS file
(-S, with and without -d).  Stored per case: the records (every positional alignformat
argument, oracle.PCR_REF_FORMAT), the text of the default output format and of a custom
format with the remaining codes.  Seeds on which the reference aborts are skipped.
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

N_CASES = 80
# the alignformat codes PCR_REF_FORMAT leaves out (the reverse-complemented amplicon %* only for -S cases)
CUSTOM_FORMAT = (r"%H|%h|%>p %<p|%>q %<q|%>Q %<Q|%>t %<t|%>T %<T|%R|%>R %<R|%>l %<l %l|%!>s|%!l|%!<e|%!>A|%!<Q"
                 r"|%I|%L|%>L %<L|%D|%a|%O|%&|%X|%0|%%|%z|\\ \t\n")


def sts_lines(rng, primers):
    """One UniSTS line per pair: id, forward, reverse, size or lo-hi, accession, chromosome,
    alternative accession, species (free text); sometimes only the first columns."""
    out, sizes = [], []
    for j in range(len(primers) // 2):
        lo = rng.choice([60, 150, 200, 350, 700])
        hi = lo + rng.choice([0, 0, 25, 300])
        sizes.append((lo, hi))
        size = str(lo) if lo == hi else "%d-%d" % (lo, hi)
        cols = ["STS%d" % (100 + j), primers[2 * j], primers[2 * j + 1], size]
        if rng.random() < 0.8:
            cols += ["G%05d" % rng.randrange(10 ** 5), rng.choice(["1", "X", "12"]), "ALT%d" % j, "Homo sapiens"]
        out.append("\t".join(cols))
    return "\n".join(out) + "\n", sizes


def make_pcr_case(seed: int, k: int, indels: bool, wc: bool, rev_comp: bool):
    """(fasta text, primers): primers[2j], primers[2j+1] = pair j as the user would give it."""
    rng = random.Random(seed)
    npairs = rng.choice([1, 2, 4])
    primers = []
    for _ in range(2 * npairs):
        p = casegen.rand_dna(rng, rng.randint(16, 24))
        if wc:
            p = list(p)
            for _ in range(rng.randint(0, 2)):
                p[rng.randrange(len(p))] = rng.choice(casegen.IUPAC)
            p = "".join(p)
        primers.append(p)

    def concrete(p):
        return "".join(rng.choice("ACGT") if c not in "ACGT" else c for c in p)

    seqs = []
    for si in range(rng.choice([1, 2, 3])):
        parts = [casegen.rand_dna(rng, rng.randint(0, 200))]
        for _ in range(rng.randint(2, 7)):
            j = rng.randrange(npairs)
            f, r = primers[2 * j], primers[2 * j + 1]
            top_r = O.reverse_comp(r) if rev_comp else r     # what the right end looks like on the top strand
            shape = rng.random()
            if shape < 0.4:
                left, right = f, top_r
            elif shape < 0.8:
                left, right = O.reverse_comp(top_r), O.reverse_comp(f)
            else:                                            # the other orientations (-a)
                cand = [f, O.reverse_comp(f), top_r, O.reverse_comp(top_r)]
                left, right = rng.choice(cand), rng.choice(cand)
            left = casegen.mutate(rng, concrete(left), rng.choice([0, 0, 0, 1, k, k + 1]), indels)
            right = casegen.mutate(rng, concrete(right), rng.choice([0, 0, 0, 1, k, k + 1]), indels)
            spacer = list(casegen.rand_dna(rng, rng.choice([0, 3, 40, 40, 150, 150, 300, 600, 1500, 2300])))
            if spacer and rng.random() < 0.3:
                a = rng.randrange(len(spacer))
                spacer[a:a + rng.randint(1, 12)] = "N" * min(12, len(spacer) - a)
            if rng.random() < 0.15:
                spacer = []
                right = ""                                   # a lone primer site
            parts += [left, "".join(spacer), right, casegen.rand_dna(rng, rng.randint(0, 120))]
        seqs.append("".join(parts))
    out = []
    for i, s in enumerate(seqs):
        out.append(">amp%d pcr case %d" % (i + 1, seed))
        for j in range(0, len(s), 70):
            out.append(s[j:j + 70])
    return "\n".join(out) + "\n", primers


def main():
    if not os.path.exists(O.PCR_REF_BIN):
        raise SystemExit("oracle/_ref/pcr_match missing: run `make -C oracle` first")
    cases, skipped, seed = [], 0, -1
    tmp = tempfile.mkdtemp()
    while len(cases) < N_CASES:
        seed += 1
        rng = random.Random(7700 + seed)
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
        use_sts = rng.random() < 0.25
        if use_sts:
            params["rev_comp"] = rev_comp = True             # -S implies -r
            params["deviation"] = rng.choice([-1, 0, 30, 400])
        fasta, primers = make_pcr_case(7700 + seed, k, indels, wc, rev_comp)
        path = os.path.join(tmp, "p%d.fasta" % seed)
        with open(path, "w", newline="") as f:
            f.write(fasta)
        case = dict(seed=seed, params=params, fasta=fasta, primers=primers)
        run = dict(params)
        if use_sts:
            case["sts"], case["sts_sizes"] = sts_lines(rng, primers)
            run["sts_path"] = os.path.join(tmp, "p%d.sts" % seed)
            with open(run["sts_path"], "w") as f:
                f.write(case["sts"])
        try:
            case["stdout_records"] = O.run_reference_pcr(path, primers, **run)
            case["stdout_default"] = O.run_reference_pcr(path, primers, fmt=None, **run)
            case["stdout_custom"] = O.run_reference_pcr(path, primers, fmt=CUSTOM_FORMAT + ("%*\\n" if use_sts else ""),
                                                        **run)
        except RuntimeError:
            skipped += 1
            continue
        cases.append(case)
    out = os.path.join(HERE, "pcr_match_golden.json")
    with open(out, "w") as f:
        json.dump(dict(generator="tests/golden/make_pcr_golden.py", skipped_reference_aborts=skipped, cases=cases),
                  f, indent=0)
    nrec = sum(len(c["stdout_records"].splitlines()) for c in cases)
    print("wrote %s: %d cases (%d from UniSTS files), %d records (%d cases with none), %d reference aborts skipped" % (
        out, len(cases), sum(1 for c in cases if "sts" in c), nrec,
        sum(1 for c in cases if not c["stdout_records"]), skipped))


if __name__ == "__main__":
    main()
