"""Seeded generators of small FASTA databases and primer sets that exercise the
edge cases of the PrimerMatch scan: planted near-matches (substitutions and
indels), hits hanging off sequence ends, hits straddling sequence boundaries,
low-complexity repeats (chains of adjacent candidates), IUPAC wildcards in
primers and text, N runs, lower-case text, short / empty sequences, CRLF.

Shared by tests/golden/make_golden.py (fixtures from the real reference),
the oracle tests and the GPU parity tests, so all of them see the same inputs.
"""
from __future__ import annotations

import random

IUPAC = "RYKMSWBDHVN"


def rand_dna(rng: random.Random, n: int, alphabet="ACGT") -> str:
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng: random.Random, s: str, nedits: int, indels: bool) -> str:
    s = list(s)
    for _ in range(nedits):
        if not s:
            break
        op = rng.choice("sid") if indels else "s"
        i = rng.randrange(len(s))
        if op == "s":
            s[i] = rng.choice([c for c in "ACGT" if c != s[i]])
        elif op == "i":
            s.insert(i, rng.choice("ACGT"))
        else:
            del s[i]
    return "".join(s)


def make_case(seed: int, nseq=3, seqlen=400, nprimers=6, plen=(12, 22), k=2, indels=True,
              wc=False, text_iupac=0.0, lower=0.0, repeats=True, line=60, crlf=False):
    """Returns (fasta_text, primers)."""
    rng = random.Random(seed)
    primers = []
    for _ in range(nprimers):
        m = rng.randint(*plen)
        p = rand_dna(rng, m)
        if wc:
            p = list(p)
            for _ in range(rng.randint(0, 3)):
                p[rng.randrange(m)] = rng.choice(IUPAC)
            p = "".join(p)
        primers.append(p)
    if repeats and nprimers >= 2:
        m = rng.randint(*plen)
        primers[0] = ("A" * m)                      # poly-A: long candidate chains
        unit = rng.choice(["AC", "AGT", "CA"])
        primers[1] = (unit * m)[:m]                 # microsatellite
    seqs = []
    for si in range(nseq):
        n = rng.randint(max(1, seqlen // 4), seqlen) if si else seqlen
        s = list(rand_dna(rng, n))
        # plant primers (or their mutated / reverse-complemented copies)
        for _ in range(rng.randint(2, 6)):
            p = rng.choice(primers)
            q = "".join(rng.choice("ACGT") if c not in "ACGT" else c for c in p)
            if rng.random() < 0.5:
                q = q[::-1].translate(str.maketrans("ACGT", "TGCA"))
            q = mutate(rng, q, rng.randint(0, k + 1), indels)
            where = rng.choice(["mid", "mid", "start", "end"])
            if where == "start":
                cut = rng.randint(0, 3)
                s[0:len(q) - cut] = list(q[cut:])
            elif where == "end":
                cut = rng.randint(0, 3)
                q2 = q[:len(q) - cut]
                if len(q2) <= len(s):
                    s[len(s) - len(q2):] = list(q2)
            else:
                if len(s) > len(q):
                    at = rng.randrange(0, len(s) - len(q))
                    s[at:at + len(q)] = list(q)
        if repeats and si == 0 and n > 120:
            s[40:40 + 45] = list("A" * 45)
            s[100:100 + 40] = list(("AC" * 40)[:40])
        for i in range(len(s)):
            r = rng.random()
            if r < text_iupac:
                s[i] = rng.choice(IUPAC)
            elif r < text_iupac + lower:
                s[i] = s[i].lower()
        seqs.append("".join(s))
    if nseq >= 3:
        seqs[1] = seqs[1][:rng.randint(1, 15)]        # a very short sequence
    nl = "\r\n" if crlf else "\n"
    out = []
    for i, s in enumerate(seqs):
        out.append(">seq%d case %d" % (i + 1, seed))
        for j in range(0, len(s), line):
            out.append(s[j:j + line])
    return nl.join(out) + nl, primers
