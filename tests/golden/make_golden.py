"""Generates tests/golden/primer_match_golden.json from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref/primer_match, which oracle/Makefile
compiles from /root/reference):

    make -C oracle && python tests/golden/make_golden.py

Each case is a seeded FASTA database + primer set from casegen.make_case and
the records the reference printed for it with
    -A "%f\\t%s\\t%e\\t%r\\t%i\\t%d\\t%S\\t%E\\t%A\\n"
(-N 5 for k > 0, i.e. the filter_bitvec index; automatic choice for k == 0).
Cases on which the reference aborts (its StreamedFastaFile asserts when an
alignment buffer reaches back to the start of the stream, fasta_io.t:713, and it
exits on a "Bogus hit" when its exact wildcard scan and exact_wc_alignment
disagree about lower-case text) are skipped: they have no reference answer.
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

N_CASES = 60


def case_params(seed: int):
    rng = random.Random(1000 + seed)
    k = rng.choice([0, 1, 2, 2, 3])
    indels = rng.random() < 0.5
    wc = rng.random() < 0.4
    tn = wc and rng.random() < 0.5
    rc = rng.random() < 0.6
    kw = dict(nseq=rng.choice([1, 3, 5]), seqlen=rng.choice([60, 400, 1500]),
              nprimers=rng.choice([1, 2, 6, 12]),
              plen=rng.choice([(8, 12), (12, 22), (20, 20), (25, 32)]), k=k, indels=indels, wc=wc,
              text_iupac=rng.choice([0, 0, 0.01, 0.05]), lower=rng.choice([0, 0, 0.02]),
              repeats=rng.random() < 0.7, line=rng.choice([60, 70, 13]), crlf=rng.random() < 0.15)
    return dict(k=k, indels=indels, wc=wc, tn=tn, rev_comp=rc), kw


def cli_args(fasta_path, primers, params, extra=()):
    """primer_match command line for a case (everything after the program name)."""
    args = ["-i", fasta_path, "-p", " ".join(primers)]
    if params["k"] > 0:
        args += ["-k" if params["indels"] else "-K", str(params["k"]), "-N", "5"]
    if params["rev_comp"]:
        args.append("-r")
    if params["wc"]:
        args.append("-W" if params["tn"] else "-w")
    return args + list(extra)


def reference_stdout(fasta_path, primers, params, extra=()):
    import subprocess
    p = subprocess.run([O.REF_BIN] + cli_args(fasta_path, primers, params, extra), capture_output=True, timeout=600)
    if p.returncode != 0:
        raise RuntimeError("reference failed")
    return p.stdout.decode("latin-1")


def main():
    if not O.ref_available():
        raise SystemExit("oracle/_ref/primer_match missing: run `make -C oracle` first")
    cases = []
    seed = 0
    tmp = tempfile.mkdtemp()
    skipped = 0
    while len(cases) < N_CASES:
        params, kw = case_params(seed)
        fasta, primers = casegen.make_case(seed, **kw)
        path = os.path.join(tmp, "c%d.fasta" % seed)
        with open(path, "w", newline="") as f:
            f.write(fasta)
        try:
            recs = O.run_reference(path, primers, **params)
        except RuntimeError:
            skipped += 1
            seed += 1
            continue
        case = dict(seed=seed, params=params, fasta=fasta, primers=primers,
                    records=[[r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index,
                              r.editdist, r.stream_start, r.stream_end, r.alignment]
                             for r in recs])
        if len(cases) % 3 == 0:
            # the tool's own text output: default alignment format, a custom one, counts
            case["stdout_default"] = reference_stdout(path, primers, params)
            case["stdout_custom"] = reference_stdout(path, primers, params,
                                                     ["-A", "%H %f %5 %3 %l %q %p %t|%T|%Q|%A %| %^ %v %* %+ %%\\n"])
            case["stdout_counts"] = reference_stdout(path, primers, params, ["-c"])
            case["stdout_counts_agg"] = reference_stdout(path, primers, params, ["-C", "%i\\t%p\\t%c\\t%C\\n", "-a"])
        cases.append(case)
        seed += 1
    # the reference's own primer lists (db/pat.txt, db/patwc.txt: BASELINE.json configs[0] and
    # configs[2]) against seeded DNA with mutated copies planted
    refdb = os.path.join(os.environ.get("PM_REFERENCE", "/root/reference"), "db")
    for name, plist, variants in (
            ("pat.txt", "pat.txt", [dict(k=0, indels=True, wc=False, tn=False, rev_comp=True),
                                    dict(k=1, indels=False, wc=False, tn=False, rev_comp=True),
                                    dict(k=2, indels=True, wc=False, tn=False, rev_comp=True)]),
            ("patwc.txt", "patwc.txt", [dict(k=0, indels=True, wc=True, tn=False, rev_comp=True),
                                        dict(k=2, indels=False, wc=True, tn=False, rev_comp=True),
                                        dict(k=1, indels=True, wc=True, tn=True, rev_comp=False)])):
        with open(os.path.join(refdb, plist)) as f:
            primers = f.read().split()
        for vi, params in enumerate(variants):
            rng = random.Random(7000 + vi)
            seqs = []
            for si in range(3):
                s = list(casegen.rand_dna(rng, 3000))
                for p in primers:
                    q = "".join(rng.choice("ACGT") if c not in "ACGT" else c for c in p)
                    if rng.random() < 0.5:
                        q = q[::-1].translate(str.maketrans("ACGT", "TGCA"))
                    q = casegen.mutate(rng, q, rng.randint(0, params["k"] + 1), params["indels"])
                    at = rng.randrange(100, len(s) - 100)
                    s[at:at + len(q)] = list(q)
                seqs.append("".join(s))
            fasta = "".join(">ref%s_%d\n%s\n" % (name, i, "\n".join(s[j:j + 70] for j in range(0, len(s), 70)))
                            for i, s in enumerate(seqs))
            path = os.path.join(tmp, "ref_%s_%d.fasta" % (name, vi))
            with open(path, "w", newline="") as f:
                f.write(fasta)
            try:
                recs = O.run_reference(path, primers, **params)
            except RuntimeError:
                skipped += 1
                continue
            cases.append(dict(seed="ref:%s:%d" % (name, vi), params=params, fasta=fasta, primers=primers,
                              records=[[r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index,
                                        r.editdist, r.stream_start, r.stream_end, r.alignment]
                                       for r in recs]))
    out = os.path.join(HERE, "primer_match_golden.json")
    with open(out, "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", reference_format=O.REF_FORMAT,
                       skipped_reference_aborts=skipped, cases=cases), f, indent=0)
    nrec = sum(len(c["records"]) for c in cases)
    print("wrote %s: %d cases, %d records, %d reference aborts skipped" % (out, len(cases), nrec, skipped))


if __name__ == "__main__":
    main()
