"""Generates tests/golden/k45_golden.json from the UNMODIFIED reference: k = 4 and 5 (-k / -K),
the error counts the bit-plane kernels serve beyond the pair index's k <= 3.

    make -C oracle && python tests/golden/make_k45_golden.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import casegen  # noqa: E402


def cases():
    out = []
    for seed, k, indels, wc, plen in ((6001, 4, False, False, (20, 28)), (6002, 4, True, False, (22, 30)),
                                      (6003, 5, False, False, (24, 32)), (6004, 5, True, False, (26, 40)),
                                      (6005, 4, True, True, (20, 26)), (6006, 5, False, True, (30, 44)),
                                      (6007, 4, False, False, (40, 64)), (6008, 5, True, False, (48, 64))):
        out.append(dict(seed=seed, params=dict(k=k, indels=indels, wc=wc, tn=False, rev_comp=True),
                        kw=dict(nseq=3, seqlen=1500, nprimers=6, plen=list(plen), k=k, indels=indels, wc=wc,
                                text_iupac=0.01 if wc else 0.0, lower=0.0, repeats=True, line=60, crlf=False)))
    return out


def main():
    import pm_oracle as O
    if not O.ref_available():
        raise SystemExit("oracle/_ref/primer_match missing: run `make -C oracle` first")
    tmp = tempfile.mkdtemp()
    res = []
    for c in cases():
        kw = dict(c["kw"], plen=tuple(c["kw"]["plen"]))
        recs = None
        for attempt in range(20):
            # (the reference aborts when an alignment buffer reaches back to the start of a streamed
            # FASTA file, fasta_io.t:713: such a seed is skipped, the next one taken)
            fasta, primers = casegen.make_case(c["seed"], **kw)
            path = os.path.join(tmp, "k45_%d.fasta" % c["seed"])
            with open(path, "w") as f:
                f.write(fasta)
            try:
                recs = O.run_reference(path, primers, **c["params"])        # -N 5 (filter_bitvec)
                break
            except RuntimeError:
                c["seed"] += 100
        if recs is None:
            raise SystemExit("no usable seed for %r" % c)
        res.append(dict(c, fasta=fasta, primers=primers,
                        records=[[r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index, r.editdist,
                                  r.stream_start, r.stream_end, r.alignment] for r in recs]))
        print("case %d k=%d: %d records" % (c["seed"], c["params"]["k"], len(recs)), flush=True)
    with open(os.path.join(HERE, "k45_golden.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_k45_golden.py", cases=res), f, indent=0)


if __name__ == "__main__":
    main()
