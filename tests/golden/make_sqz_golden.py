"""Generates tests/golden/sqz_golden.json from the UNMODIFIED reference tools.

Seeds on which the reference's StreamedFastaFile aborts (fasta_io.t:713, see
make_golden.py) are skipped.  Run in the build container (needs oracle/_ref/compress_seq and primer_match,
compiled from /root/reference by oracle/Makefile):

    make -C oracle && python tests/golden/make_sqz_golden.py

For each seeded FASTA database: `compress_seq -n true -z true` writes the
.tbl / .tbz character tables, the .sqn file (one table index per character) and the
.sqz bit stream (compress_seq.cc:745-870); the fixture keeps them (base64) with the length of the stream (size of the .sqn file), and the records
primer_match prints when it reads the database through the compressed files
(-D 4, select.t:94-106) and through the normalised ones (-D 3) -- checked here to
equal what it prints on the FASTA file itself.
"""
from __future__ import annotations

import base64
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import casegen  # noqa: E402
import pm_oracle as O  # noqa: E402

CSQ = os.path.join(ROOT, "oracle", "_ref", "compress_seq")


def main():
    if not (O.ref_available() and os.path.exists(CSQ)):
        raise SystemExit("oracle/_ref/{primer_match,compress_seq} missing: run `make -C oracle` first")
    cases = []
    seed = -1
    while len(cases) < 12:
        seed += 1
        rng = random.Random(4000 + seed)
        k = rng.choice([0, 1, 2])
        indels = rng.random() < 0.5
        params = dict(k=k, indels=indels, wc=False, tn=False, rev_comp=True)
        kw = dict(nseq=rng.choice([1, 2, 4]), seqlen=rng.choice([50, 333, 1200]), nprimers=rng.choice([2, 5]),
                  plen=rng.choice([(12, 20), (20, 20)]), k=k, indels=indels, wc=False,
                  text_iupac=rng.choice([0, 0.02, 0.3]), lower=0, repeats=True, line=rng.choice([60, 17]), crlf=False)
        fasta, primers = casegen.make_case(9000 + seed, **kw)
        tmp = tempfile.mkdtemp()
        try:
            path = os.path.join(tmp, "db.fasta")
            with open(path, "w", newline="") as f:
                f.write(fasta)
            try:
                plain = O.run_reference(path, primers, **params)       # streamed FASTA
            except RuntimeError:                                        # fasta_io.t:713 abort: no answer
                continue
            subprocess.run([CSQ, "-i", path, "-n", "true", "-z", "true"], check=True, capture_output=True)
            n_chars = os.path.getsize(path + ".sqn")                    # one byte per stream character
            via_sqn = O.run_reference(path, primers, extra=["-D", "3"], **params)   # Normalized<>
            with open(path + ".sqn", "rb") as f:
                sqn = f.read()
            with open(path + ".tbl", "rb") as f:
                tbl = f.read()
            os.unlink(path + ".sqn")                                    # leave only the compressed form
            via_sqz = O.run_reference(path, primers, extra=["-D", "4"], **params)
            if plain != via_sqz or plain != via_sqn:
                raise SystemExit("reference disagrees with itself on seed %d" % seed)
            with open(path + ".tbz", "rb") as f:
                tbz = f.read()
            with open(path + ".sqz", "rb") as f:
                sqz = f.read()
            cases.append(dict(seed=seed, params=params, fasta=fasta, primers=primers, n_chars=n_chars,
                              tbz=base64.b64encode(tbz).decode(), sqz=base64.b64encode(sqz).decode(),
                              tbl=base64.b64encode(tbl).decode(), sqn=base64.b64encode(sqn).decode(),
                              records=[[r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index,
                                        r.editdist, r.stream_start, r.stream_end, r.alignment] for r in via_sqz]))
        finally:
            shutil.rmtree(tmp)
    out = os.path.join(HERE, "sqz_golden.json")
    with open(out, "w") as f:
        json.dump(dict(generator="tests/golden/make_sqz_golden.py", cases=cases), f, indent=0)
    print("wrote %s: %d cases, %d records, table sizes %s" % (
        out, len(cases), sum(len(c["records"]) for c in cases),
        sorted({len(base64.b64decode(c["tbz"])) for c in cases})))


if __name__ == "__main__":
    main()
