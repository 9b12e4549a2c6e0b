"""Generates tests/golden/primer_match_sts_golden.json from the UNMODIFIED reference.

    make -C oracle && python tests/golden/make_sts_golden.py

primer_match -S: the primers are the forward and reverse primers of every line of a UniSTS
file (primer_match.cc:901-916) and the output formats gain the codes %I %L %a %O %& %X
(alignformat, primer_match.cc:500-521; countformat, 791-812).  Stored per case: the
reference's text for an alignment format and a count format made of those codes.
"""
from __future__ import annotations

import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import pm_oracle as O  # noqa: E402
from make_pcr_golden import make_pcr_case, sts_lines  # noqa: E402

ALIGN_FORMAT = r"%f\t%s\t%e\t%r\t%i\t%d\t%S\t%E\t%A\t%I|%L|%a|%O|%&|%X|%U|%F|%n|%D\n"
COUNT_FORMAT = r"%i %r %c ( %C ) %I|%L|%a|%O|%&|%X\n"


def sts_cli_args(fasta_path, sts_path, params):
    args = ["-i", fasta_path, "-S", sts_path]
    if params["k"] > 0:
        args += ["-k" if params["indels"] else "-K", str(params["k"]), "-N", "5"]
    if params["rev_comp"]:
        args.append("-r")
    return args


def main():
    if not O.ref_available():
        raise SystemExit("oracle/_ref/primer_match missing: run `make -C oracle` first")
    cases, skipped, seed = [], 0, -1
    tmp = tempfile.mkdtemp()
    while len(cases) < 10:
        seed += 1
        rng = random.Random(5100 + seed)
        params = dict(k=rng.choice([0, 1, 2]), indels=rng.random() < 0.5, rev_comp=rng.random() < 0.7, wc=False, tn=False)
        fasta, primers = make_pcr_case(5100 + seed, params["k"], params["indels"], False, True)
        sts, _ = sts_lines(rng, primers)
        fa, sp = os.path.join(tmp, "s%d.fasta" % seed), os.path.join(tmp, "s%d.sts" % seed)
        with open(fa, "w", newline="") as f:
            f.write(fasta)
        with open(sp, "w") as f:
            f.write(sts)
        out = {}
        for key, extra in (("stdout_align", ["-A", ALIGN_FORMAT]), ("stdout_counts", ["-C", COUNT_FORMAT]),
                           ("stdout_counts_agg", ["-C", COUNT_FORMAT, "-a"])):
            p = subprocess.run([O.REF_BIN] + sts_cli_args(fa, sp, params) + extra, capture_output=True, timeout=600)
            if p.returncode != 0:
                out = None
                break
            out[key] = p.stdout.decode("latin-1")
        if out is None:
            skipped += 1
            continue
        cases.append(dict(seed=seed, params=params, fasta=fasta, primers=primers, sts=sts, **out))
    path = os.path.join(HERE, "primer_match_sts_golden.json")
    with open(path, "w") as f:
        json.dump(dict(generator="tests/golden/make_sts_golden.py", skipped_reference_aborts=skipped, cases=cases), f,
                  indent=0)
    print("wrote %s: %d cases, %d records, %d reference aborts skipped" % (
        path, len(cases), sum(len(c["stdout_align"].splitlines()) for c in cases), skipped))


if __name__ == "__main__":
    main()
