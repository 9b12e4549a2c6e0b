/*
 * pm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference PrimerMatch database scan, used
 * as the parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product (sequence-alignment-tools_b200/) never does.
 *
 * Parity status: PINNED.  The reference ships no golden outputs for this path
 * (its db/pri_non_hum_nt.fasta is a missing large blob and testscript.sh only
 * checks that its own backends agree with each other), so the pin is the
 * reference itself: oracle/_ref/primer_match is the unmodified reference
 * compiled from /root/reference by oracle/Makefile, tests/golden/make_golden.py
 * runs it on seeded inputs and commits the outputs as fixtures, and
 * tests/test_oracle_golden.py checks every function below against them.
 *
 * Each function cites the reference file:line it restates.  Restatement
 * choices that are NOT literal:
 *   - the reference concatenates all patterns into one long bit string spread
 *     over 64-bit words (shift_and_inexact.cc:95-178).  Patterns do not
 *     interact: the only cross-pattern term is the shift carry into a
 *     pattern's first bit, and that bit is always OR-ed with the start mask
 *     s_ anyway.  The oracle therefore keeps one 64-bit word per pattern per
 *     level (pattern length <= 64), which is what the CUDA kernel does too.
 *   - batching (find_patterns' minka / report_interval hand-offs,
 *     filter_bitvec.cc:141-234) is transparent to the hit set: a chain is only
 *     processed once the scan frontier is 2k+1 past its last candidate, i.e.
 *     when no later candidate can still join it.  The oracle scans the whole
 *     stream, then chains.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ------------------------------------------------------------------------- */
/* alignment codes: alignment_code.h:5-36                                     */
enum {
  AC_NONE = 0, AC_EQUAL = 1, AC_WC_EQUAL = 2, AC_SUBST = 3, AC_INSERT = 4,
  AC_DELETE = 5, AC_CONSTRAINT = 6, AC_END = 7
};
enum {
  M_EQUAL = 2, M_WC = 4, M_SUBST = 8, M_INSERT = 16, M_DELETE = 32,
  M_CONSTRAINT = 64, M_END = 128
};

/* ------------------------------------------------------------------------- */
/* IUPAC compatibility: util.cc:121-185 (iupac_compatible).  Row = pattern    */
/* character, entries = text characters it is compatible with.               */
static const char *iupac_row_upper(int w) {
  switch (w) {
  case 'A': return "ARMWDHVN";
  case 'B': return "GTUCYKSBN";
  case 'C': return "CYMSBHVN";
  case 'D': return "GATURWKDN";
  case 'G': return "GRKSBDVN";
  case 'H': return "ACTUMYWHN";
  case 'K': return "GTKBDN";
  case 'M': return "ACMHVN";
  case 'N': return "ACGTURYKMSWVDHVN";
  case 'R': return "GARDVN";
  case 'S': return "GCSBVN";
  case 'T': return "TUYKWVDHN";
  case 'U': return "UTYKWVDHN";
  case 'V': return "GCARSMVN";
  case 'W': return "ATUWDHN";
  case 'Y': return "TUCYBHN";
  case 'X': return "MRWSYKVHDBXN";
  default: return 0;
  }
}

/* compat[w][c] for 7-bit w,c; lower-case rows are the lower-cased upper rows
 * (util.cc:143-159). */
static unsigned char g_compat[128][128];
static unsigned char g_hasrow[128];
static int g_compat_init = 0;

static void compat_init(void) {
  if (g_compat_init) return;
  memset(g_compat, 0, sizeof g_compat);
  memset(g_hasrow, 0, sizeof g_hasrow);
  for (int w = 'A'; w <= 'Z'; w++) {
    const char *row = iupac_row_upper(w);
    if (!row) continue;
    g_hasrow[w] = 1;
    g_hasrow[tolower(w)] = 1;
    for (const char *p = row; *p; p++) {
      g_compat[w][(int)*p] = 1;
      g_compat[tolower(w)][tolower(*p)] = 1;
    }
  }
  g_compat_init = 1;
}

/* Does pattern character pc match text character tc in the *scanner's* mask
 * table?  shift_and_inexact.cc:126-143 (and shift_and.cc:107-119):
 *   wc && pattern char has a compat row -> every row member except 'N'
 *   (upper-case N only) unless textn;  otherwise literal equality.          */
int pmo_scan_match(int pc, int tc, int wc, int tn) {
  compat_init();
  pc &= 0xff; tc &= 0xff;
  if (wc && pc < 128 && g_hasrow[pc]) {
    if (tc >= 128) return 0;
    if (!g_compat[pc][tc]) return 0;
    if (tc == 'N' && !tn) return 0;
    return 1;
  }
  return pc == tc;
}

/* ------------------------------------------------------------------------- */
/* FASTA -> normalised character stream.                                      */
/* StreamedFastaFile::getch, fasta_io.t:478-629, with eos_start=true          */
/* (fasta_io.t:20).  Output stream: eos, seq1, eos, seq2, ... seqN, eos.      */
/* seq_key[i]  = stream index of the first character of sequence i            */
/*               (the key the header is inserted at, fasta_io.t:535-550)      */
/* hdr_off/len = byte range of the header text (after '>') in raw            */
/* Returns stream length; *nseq_out = number of headers registered.          */
long pmo_normalize_fasta(const uint8_t *raw, long n, int uc, uint8_t eos,
                         uint8_t *out, long *seq_key, long *hdr_off,
                         long *hdr_len, long maxseq, long *nseq_out) {
  long rp = 0;      /* raw position (SortSeqFastaFile::pos) */
  long pos = 0;     /* _pos */
  long nseq = 0;
  long lastkey = 0; /* this->lastkey(), fasta_io.t:217-223 */
  int eof_emitted = 0;
  for (;;) {
    /* fasta_io.t:479-483 */
    if (rp >= n && !eof_emitted) {
      eof_emitted = 1;
      out[pos++] = eos;
      break;
    }
    /* fasta_io.t:484-486 */
    if (pos == 0) { out[pos++] = eos; continue; }
    uint8_t ch = raw[rp++];
    if (ch != '\n' && ch != '\r' && ch != ' ' && ch != '>') {
      out[pos++] = uc ? (uint8_t)toupper(ch) : ch;   /* fasta_io.t:489-496 */
      continue;
    }
    /* fasta_io.t:500-510: swallow white space; EOF inside it ends the stream */
    int hit_eof = 0;
    while (ch == '\n' || ch == '\r' || ch == ' ') {
      if (rp >= n) { hit_eof = 1; break; }
      ch = raw[rp++];
    }
    if (hit_eof) { eof_emitted = 1; out[pos++] = eos; break; }
    if (ch == '>') {
      /* fasta_io.t:525-534: header runs to end of line */
      long hstart = rp;
      uint8_t c2 = 0;
      while (rp < n) { c2 = raw[rp++]; if (c2 == '\n' || c2 == '\r') break; }
      long hend = rp;
      if (c2 == '\r' && rp < n) rp++;          /* consume the '\n' of CRLF */
      int first = (pos == 1);
      if (!first) pos++;                        /* fasta_io.t:535-537 */
      if (lastkey < pos || first) {             /* fasta_io.t:546-551 */
        if (nseq < maxseq) {
          seq_key[nseq] = pos;
          hdr_off[nseq] = hstart;
          hdr_len[nseq] = hend - hstart - 1;
        }
        nseq++;
        lastkey = pos;
      }
      if (first) {
        /* fasta_io.t:575-583: the character after the first header line is
         * returned verbatim, whatever it is */
        if (rp < n) {
          uint8_t c3 = raw[rp++];
          out[pos++] = uc ? (uint8_t)toupper(c3) : c3;
        }
        /* (reading past a file that ends right after its first header is
         * undefined in the reference; the oracle stops cleanly) */
      } else {
        out[pos - 1] = eos;                     /* fasta_io.t:585-586 */
      }
    } else {
      out[pos++] = uc ? (uint8_t)toupper(ch) : ch;   /* fasta_io.t:618-625 */
    }
  }
  *nseq_out = nseq;
  return pos;
}

/* ------------------------------------------------------------------------- */
/* Candidate record: pos is CharacterProducer::pos() after the last matched   */
/* character was read, i.e. (stream index of that character) + 1.             */
typedef struct { int64_t pos; int32_t pid; int32_t val; } pmo_hit;

/* ------------------------------------------------------------------------- */
/* k-error bit-vector scan: shift_and_inexact::computeu + find_patterns,      */
/* shift_and_inexact.cc:88-178 and 237-352.  One 64-bit word per pattern per  */
/* level (see header).  pats[p] are the pattern strings, pid reported is p+1  */
/* (ids are assigned 1.. in add order, pattern_match.h:95-99).                */
/* Candidates are emitted in stream order, and within one position in         */
/* pattern order (the reference walks words then bits, lines 307-326).        */
/* Returns the number of candidates (may exceed cap; only cap are stored).    */
long pmo_wm_scan(const uint8_t *T, long L, int npat, const char *const *pats,
                 int k, int indels, int wc, int tn, uint8_t eos,
                 pmo_hit *out, long cap) {
  compat_init();
  if (k > 15) return -1;
  /* per-pattern masks u[c], computeu lines 121-146 */
  uint64_t *U = (uint64_t *)calloc((size_t)npat * 256, sizeof(uint64_t));
  uint64_t *last = (uint64_t *)calloc((size_t)npat, sizeof(uint64_t));
  uint64_t *M = (uint64_t *)calloc((size_t)npat * (size_t)(k + 1), sizeof(uint64_t));
  for (int p = 0; p < npat; p++) {
    int m = (int)strlen(pats[p]);
    if (m < 1 || m > 64) { free(U); free(last); free(M); return -1; }
    for (int i = 0; i < m; i++)
      for (int c = 0; c < 256; c++)
        if (pmo_scan_match((unsigned char)pats[p][i], c, wc, tn))
          U[(size_t)p * 256 + c] |= 1ULL << i;
    last[p] = 1ULL << (m - 1);
    /* initial state, computeu lines 147-149: level l has bits 0..l-1 set */
    for (int i = 0; i < m; i++)
      for (int l = i + 1; l <= k; l++)
        M[(size_t)p * (k + 1) + l] |= 1ULL << i;
  }
  long n = 0;
  const uint64_t S = 1ULL;          /* s_: first bit of the pattern */
  for (long t = 0; t < L; t++) {
    uint8_t ch = T[t];
    for (int p = 0; p < npat; p++) {
      uint64_t *m_ = M + (size_t)p * (k + 1);
      uint64_t uch = U[(size_t)p * 256 + ch];
      /* lines 253-267 */
      uint64_t m0 = m_[0];
      uint64_t m1 = (m_[0] << 1) | S;
      m_[0] = m1 & uch;
      if (indels) m1 |= m0;
      /* lines 268-303 */
      for (int l = 1; l <= k; l++) {
        uint64_t m2 = m_[l];
        uint64_t m3 = (m_[l] << 1) | S;
        m_[l] = m3 & uch;
        if (ch != eos) {
          m_[l] |= m1;
          if (indels) {
            m_[l] |= (m_[l - 1] << 1) | S;     /* NEW level l-1, line 283/294 */
            m_[l] |= m_[l - 1];                /* line 284/295 */
          }
          m1 = m3;
          if (indels) m1 |= m2;
        }
        /* when ch == eos m1 is left untouched (line 277 guards the whole
         * block); every level is zero anyway because uch == 0 */
      }
      /* lines 305-326 */
      if (m_[k] & last[p]) {
        int kk = k - 1;
        while (kk >= 0 && (m_[kk] & last[p])) kk--;
        kk++;
        if (n < cap) { out[n].pos = t + 1; out[n].pid = p + 1; out[n].val = kk; }
        n++;
      }
    }
  }
  free(U); free(last); free(M);
  return n;
}

/* ------------------------------------------------------------------------- */
/* Exact multi-pattern scan: every (end position, pattern) occurrence.        */
/* keyword_tree<>::find_patterns (keyword_tree.t:427-486) when !wc, and       */
/* shift_and::find_patterns (shift_and.cc:204-256) when wc; both enumerate    */
/* all occurrences, the second under the scanner's wildcard relation.         */
long pmo_exact_scan(const uint8_t *T, long L, int npat, const char *const *pats,
                    int wc, int tn, pmo_hit *out, long cap) {
  compat_init();
  long n = 0;
  int *len = (int *)malloc(sizeof(int) * (size_t)npat);
  for (int p = 0; p < npat; p++) len[p] = (int)strlen(pats[p]);
  for (long t = 0; t < L; t++) {
    for (int p = 0; p < npat; p++) {
      int m = len[p];
      if (t + 1 < m) continue;
      int ok = 1;
      for (int i = 0; i < m && ok; i++)
        ok = pmo_scan_match((unsigned char)pats[p][i], T[t - m + 1 + i], wc, tn);
      if (ok) {
        if (n < cap) { out[n].pos = t + 1; out[n].pid = p + 1; out[n].val = 0; }
        n++;
      }
    }
  }
  free(len);
  return n;
}

/* ------------------------------------------------------------------------- */
/* editdist_alignment::align, pattern_alignment.cc:117-705 (dna_mut=false,    */
/* trans=false).  e1 = end() (poslb), e2 = end2_ (posub).                     */
/* Outputs: *start,*end in stream coordinates, *val = bestval, codes[] =      */
/* alignment codes in pattern order, *ncodes.  Return value = the function's  */
/* bool.  *editdist = pattern_alignment::editdist() (pattern_alignment.h:108) */
/* or -1 for MAXINT (constraint violation).  When the function bails out      */
/* early (lines 425-436, 482-490) start/end/val are left untouched, as in the */
/* reference, and codes = {constraint violation}.                             */
static inline uint8_t stream_at(const uint8_t *T, long L, long i) {
  /* characters past the end of the stream are never compared by the DP
   * (only buffer[0..buflen) is read); guard anyway */
  return (i >= 0 && i < L) ? T[i] : 0;
}

int pmo_editdist_align(const uint8_t *T, long L, const char *pat,
                       int64_t e1, int64_t e2, int k, uint8_t eos, int wc, int tn,
                       int indels, int lconst, int rconst,
                       int64_t *start, int64_t *end, int *val, int *editdist,
                       uint8_t *codes, int *ncodes) {
  compat_init();
  const int m = (int)strlen(pat);
  const int cvp = 5 * k + 1;                         /* line 131 */
  int64_t textstart = 0;
  if (e1 > (int64_t)m + k) textstart = e1 - m - k;   /* lines 137-139 */
  const long buflen = (long)(e2 - textstart);        /* line 148 */
  const int kk = indels ? k : 0;
  const long R = (long)(e2 - e1);
  *ncodes = 0; *editdist = -1;
  if (buflen < 0) return 0;

  const long W = buflen + 1;
  unsigned *dp = (unsigned *)malloc(sizeof(unsigned) * (size_t)W * (size_t)(m + 1));
  int *best = (int *)malloc(sizeof(int) * (size_t)W * (size_t)(m + 1));
  /* poison so that any read outside the band the reference fills is loud */
  for (long i = 0; i < W * (m + 1); i++) { dp[i] = 0x3fffffff; best[i] = 0; }
#define IDX(p, t) ((long)(p) * W + (long)(t))
#define BUF(t) stream_at(T, L, textstart + buflen - (t))   /* buffer[buflen-t] */
#define PAT(p) ((uint8_t)pat[m - (p)])

  int lbexact = 0, rbexact = m + 1;                  /* lines 230-233 */
  if (lconst > 0) rbexact = m + 1 - lconst;
  if (rconst > 0) lbexact = rconst;

  dp[IDX(0, 0)] = 0; best[IDX(0, 0)] = M_END;        /* lines 242-243 */
  long ub = kk; if (ub > m) ub = m;                  /* lines 247-250 */
  for (long p = 1; p <= ub; p++) {                   /* lines 253-268 */
    if (!indels || p < lbexact || p >= rbexact || PAT(p) == eos) {
      dp[IDX(p, 0)] = (unsigned)cvp; best[IDX(p, 0)] = M_CONSTRAINT;
    } else {
      dp[IDX(p, 0)] = dp[IDX(p - 1, 0)] + 1; best[IDX(p, 0)] = M_DELETE;
    }
  }
  ub = R + kk; if (buflen < ub) ub = buflen;         /* lines 269-272 */
  for (long t = 1; t <= ub; t++) {                   /* lines 276-294 */
    if (t <= R) {
      dp[IDX(0, t)] = 0; best[IDX(0, t)] = M_END;
    } else if (!indels || lbexact > 0) {
      dp[IDX(0, t)] = (unsigned)cvp; best[IDX(0, t)] = M_CONSTRAINT;
    } else {
      dp[IDX(0, t)] = dp[IDX(0, t - 1)] + 1; best[IDX(0, t)] = M_INSERT;
    }
  }
  int rc = -2;   /* -2: still running */
  for (long p = 1; p <= m && rc == -2; p++) {        /* lines 296-437 */
    long lb = p - kk; if (lb < 1) lb = 1;
    ub = p + R + kk; if (buflen < ub) ub = buflen;
    int bestscorerow = cvp;
    for (long t = lb; t <= ub; t++) {
      unsigned v, v1; int ac;
      uint8_t tc = BUF(t), pc = PAT(p);
      if (tc == pc) {
        v = dp[IDX(p - 1, t - 1)]; ac = M_EQUAL;
      } else if (wc && pc < 128 && tc < 128 && g_compat[pc][tc] && (tc != 'N' || tn)) {
        v = dp[IDX(p - 1, t - 1)]; ac = M_WC;
      } else if (tc == eos || pc == eos || p <= lbexact || p >= rbexact) {
        v = (unsigned)cvp; ac = M_CONSTRAINT;
      } else {
        v = dp[IDX(p - 1, t - 1)] + 1; ac = M_SUBST;
      }
      /* insertion, lines 366-391 */
      if (tc == eos || pc == eos || !indels || t <= lb || p < lbexact || p >= rbexact) {
        v1 = (unsigned)cvp;
        if (v1 < v) { v = v1; ac = M_CONSTRAINT; }
      } else {
        v1 = dp[IDX(p, t - 1)] + 1;
        if (v1 < v) { v = v1; ac = M_INSERT; } else if (v1 == v) ac |= M_INSERT;
      }
      /* deletion, lines 392-416 */
      if (!indels || pc == eos || t >= ub || p <= lbexact || p >= rbexact) {
        v1 = (unsigned)cvp;
        if (v1 < v) { v = v1; ac = M_CONSTRAINT; }
      } else {
        v1 = dp[IDX(p - 1, t)] + 1;
        if (v1 < v) { v = v1; ac = M_DELETE; } else if (v1 == v) ac |= M_DELETE;
      }
      dp[IDX(p, t)] = v; best[IDX(p, t)] = ac;
      if ((int)v < bestscorerow) bestscorerow = (int)v;
    }
    if (bestscorerow > k) rc = 0;                    /* lines 425-436 */
  }
  if (rc == 0) {
    codes[0] = AC_CONSTRAINT; *ncodes = 1; *editdist = -1;
    free(dp); free(best);
    return 0;
  }

  long beststart = m - kk;                           /* lines 443-449 */
  if (beststart > buflen) beststart = buflen;
  if (beststart < 0) beststart = 0;
  int bestval = (int)dp[IDX(m, beststart)];
  ub = m + R + kk; if (buflen < ub) ub = buflen;
  for (long kx = beststart + 1; kx <= ub; kx++) {    /* lines 461-475 */
    if ((int)dp[IDX(m, kx)] < bestval ||
        ((int)dp[IDX(m, kx)] <= bestval &&
         (best[IDX(m, kx)] & (M_EQUAL | M_WC | M_SUBST)))) {
      bestval = (int)dp[IDX(m, kx)];
      beststart = kx;
    }
  }
  long p = m, t = beststart;
  if (t < p - kk || t > p + kk + R) {                /* lines 482-490 */
    codes[0] = AC_CONSTRAINT; *ncodes = 1; *editdist = -1;
    free(dp); free(best);
    return 0;
  }
  int lastac = AC_NONE;                              /* lines 509-590 */
  int stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int nc = 0;
  while (!(best[IDX(p, t)] & M_END)) {
    int ac = best[IDX(p, t)];
    int match = ac & (M_EQUAL | M_WC | M_SUBST);
    int wcb = ac & M_WC, sub = ac & M_SUBST, ins = ac & M_INSERT, del = ac & M_DELETE;
    if (match && !((lastac == AC_INSERT && ins) || (lastac == AC_DELETE && del) ||
                   (lastac == AC_WC_EQUAL && !wcb && (ins || del)))) {
      p--; t--;
      if ((ac & M_EQUAL) && !((lastac == AC_WC_EQUAL && wcb) || (lastac == AC_SUBST && sub)))
        lastac = AC_EQUAL;
      else if (wcb) lastac = AC_WC_EQUAL;
      else if (sub) lastac = AC_SUBST;
    } else if (del) { p--; lastac = AC_DELETE; }
    else if (ins) { t--; lastac = AC_INSERT; }
    else if (ac & M_CONSTRAINT) { p = 0; t = 0; lastac = AC_CONSTRAINT; }
    else { /* assert(0) in the reference */ free(dp); free(best); return -1; }
    stats[lastac]++;
    codes[nc++] = (uint8_t)lastac;
  }
  *ncodes = nc;
  long endt = t;
  *start = e2 - beststart;                           /* line 608 */
  *end = *start + (beststart - endt);                /* lines 603-609 */
  *val = bestval;
  if (stats[AC_CONSTRAINT] > 0) *editdist = -1;      /* pattern_alignment.h:108-123 */
  else *editdist = stats[AC_SUBST] + stats[AC_INSERT] + stats[AC_DELETE];
  free(dp); free(best);
#undef IDX
#undef BUF
#undef PAT
  return bestval <= k;
}

/* ------------------------------------------------------------------------- */
/* filter_bitvec::find_patterns, filter_bitvec.cc:126-236: candidates ->      */
/* per-pattern chains (consecutive same-pattern candidates no more than 2k+1  */
/* apart) -> one yes/no alignment over [first,last] per chain.                */
/* cand must be sorted by position (the reference sorts with l.normalize());  */
/* the order of equal positions does not affect chains, which are per-pid.    */
/* lconst/rconst: per-pattern exact_start_bases/exact_end_bases (index pid-1) */
/* or NULL.  Output hits in chain-processing order: (end, pid, val).          */
long pmo_filter_chains(const uint8_t *T, long L, int npat, const char *const *pats,
                       const pmo_hit *cand, long ncand, int k, uint8_t eos,
                       int wc, int tn, int indels, const int *lconst, const int *rconst,
                       pmo_hit *out, long cap) {
  (void)npat;
  uint8_t *done = (uint8_t *)calloc((size_t)(ncand > 0 ? ncand : 1), 1);
  uint8_t codes[256];
  long n = 0;
  for (long i = 0; i < ncand; i++) {
    if (done[i]) continue;
    int pid = cand[i].pid;
    int64_t firstpos = cand[i].pos, pos = firstpos;
    done[i] = 1;
    for (long j = i + 1; j < ncand && (done[j] || cand[j].pos <= pos + 2 * k + 1); j++) {
      /* zeroed keys (done) satisfy key <= pos+2k+1 and are skipped by the pid
       * test only if they belong to another pattern; a done entry of the same
       * pid cannot follow an undone one of that pid (chains are greedy) */
      if (!done[j] && cand[j].pid == pid) { pos = cand[j].pos; done[j] = 1; }
    }
    int64_t s = 0, e = 0; int val = 0, ed = 0, nc = 0;
    int ok = pmo_editdist_align(T, L, pats[pid - 1], firstpos, pos, k, eos, wc, tn, indels,
                                lconst ? lconst[pid - 1] : 0, rconst ? rconst[pid - 1] : 0,
                                &s, &e, &val, &ed, codes, &nc);
    if (ok == 1) {
      if (n < cap) { out[n].pos = e; out[n].pid = pid; out[n].val = val; }
      n++;
    }
  }
  free(done);
  return n;
}

/* ------------------------------------------------------------------------- */
/* exact_wc_alignment::align, pattern_alignment.cc:70-93: codes for the       */
/* wildcard exact path; returns the number of substitutions (editdist).       */
int pmo_exact_wc_align(const uint8_t *T, long L, const char *pat, int64_t end,
                       int tn, uint8_t *codes) {
  compat_init();
  int m = (int)strlen(pat), sub = 0;
  for (int i = 0; i < m; i++) {
    uint8_t tc = stream_at(T, L, end - m + i), pc = (uint8_t)pat[i];
    if (tc == pc) codes[i] = AC_EQUAL;
    else if (tc < 128 && pc < 128 && g_compat[tc][pc] && (tn || tc != 'N')) codes[i] = AC_WC_EQUAL;
    else { codes[i] = AC_SUBST; sub++; }
  }
  return sub;
}

/* ------------------------------------------------------------------------- */
/* primer_alignment::global_align, primer_alignment.cc:9-462, as exact_halves */
/* uses it: yesno_ = true (no traceback), dna_mut_ = false.  text[0..textlen) */
/* is aligned against pat from its left end (dirn > 0) or from its right end  */
/* (dirn < 0) inside a band of +-maxdist (0 without indels).  lmatch / rmatch */
/* are the reference's `unsigned int` parameters: exact_halves passes         */
/* lmatch_ - pattern1.length() computed in unsigned arithmetic, and the       */
/* result lands in the int variables lbexact / rbexact (lines 43-52).         */
/* Line 97 of the reference reads `patch` before it is assigned; the          */
/* comparison is taken as false here (the pins decide: goldens made with the  */
/* reference's automatic index choice for k = 1).                             */
/* Returns the function's bool; *matchlen = bestpos, *value = bestval.        */
#define PA_M 1 /* equal / wildcard equal / substitution: the diagonal move */
#define PA_I 2
#define PA_D 4
#define PA_V 8
int pmo_primer_global_align(const uint8_t *text, unsigned textlen, const char *pat, int dirn,
                            unsigned lmatch, unsigned rmatch, int maxdist, uint8_t eos, int wc,
                            int tn, int indels, int *matchlen, int *value) {
  compat_init();
  const unsigned patlen = (unsigned)strlen(pat);
  const unsigned textlenp1 = textlen + 1, patlenp1 = patlen + 1;
  unsigned *dp = (unsigned *)malloc(sizeof(unsigned) * (size_t)textlenp1 * patlenp1);
  unsigned *best = (unsigned *)malloc(sizeof(unsigned) * (size_t)textlenp1 * patlenp1);
  for (size_t i = 0; i < (size_t)textlenp1 * patlenp1; i++) { dp[i] = 0x3fffffff; best[i] = 0; }
#define PIDX(p, t) ((size_t)(p) * textlenp1 + (size_t)(t))
  int lbexact = 0, rbexact = (int)patlenp1;
  const int cv = 5 * maxdist + 1;
  if (dirn < 0) {
    if (lmatch > 0) rbexact = (int)(patlenp1 - lmatch);
    if (rmatch > 0) lbexact = (int)rmatch;
  } else {
    if (lmatch > 0) lbexact = (int)lmatch;
    if (rmatch > 0) rbexact = (int)(patlenp1 - rmatch);
  }
  const int kk = indels ? maxdist : 0;
  int p, t, lb, ub, ret = 0;
  dp[0] = 0;
  ub = kk;
  if (patlen < (unsigned)ub) ub = (int)patlen;
  for (p = 1; p <= ub; p++) {
    if (!indels || p < lbexact || p >= rbexact) { dp[PIDX(p, 0)] = (unsigned)cv; best[PIDX(p, 0)] = PA_V; }
    else { dp[PIDX(p, 0)] = dp[PIDX(p - 1, 0)] + 1; best[PIDX(p, 0)] = PA_D; }
  }
  ub = kk;
  if (textlen < (unsigned)ub) ub = (int)textlen;
  for (t = 1; t <= ub; t++) {
    const uint8_t textch = dirn > 0 ? text[t - 1] : text[textlen - t];
    if (!indels || 0 < lbexact || 0 >= rbexact || textch == eos) { dp[PIDX(0, t)] = (unsigned)cv; best[PIDX(0, t)] = PA_V; }
    else { dp[PIDX(0, t)] = dp[PIDX(0, t - 1)] + 1; best[PIDX(0, t)] = PA_I; }
  }
  for (p = 1; p <= (int)patlen; p++) {
    lb = p - kk;
    if (lb < 1) lb = 1;
    ub = p + kk;
    if (textlen < (unsigned)ub) ub = (int)textlen;
    int bestscorerow = cv;
    for (t = lb; t <= ub; t++) {
      uint8_t textch, patch;
      if (dirn > 0) { textch = text[t - 1]; patch = (uint8_t)pat[p - 1]; }
      else { textch = text[textlen - t]; patch = (uint8_t)pat[patlen - p]; }
      int v, v1;
      unsigned ac, ac1;
      if (textch == patch) { v = (int)dp[PIDX(p - 1, t - 1)]; ac = PA_M; }
      else if (wc && textch < 128 && patch < 128 && g_compat[textch][patch] && (tn || textch != 'N')) {
        v = (int)dp[PIDX(p - 1, t - 1)]; ac = PA_M;
      } else if (textch == eos || patch == eos || p <= lbexact || p >= rbexact) { v = cv; ac = PA_V; }
      else { v = (int)dp[PIDX(p - 1, t - 1)] + 1; ac = PA_M; }
      if (textch == eos || patch == eos || !indels || t <= lb || p < lbexact || p >= rbexact) { v1 = cv; ac1 = PA_V; }
      else { v1 = (int)dp[PIDX(p, t - 1)] + 1; ac1 = PA_I; }
      if (v1 < v) { v = v1; ac = ac1; } else if (v1 == v) ac |= ac1;
      if (!indels || t >= ub || p <= lbexact || p >= rbexact) { v1 = cv; ac1 = PA_V; }
      else { v1 = (int)dp[PIDX(p - 1, t)] + 1; ac1 = PA_D; }
      if (v1 < v) { v = v1; ac = ac1; } else if (v1 == v) ac |= ac1;
      dp[PIDX(p, t)] = (unsigned)v;
      best[PIDX(p, t)] = ac;
      if (bestscorerow > v) bestscorerow = v;
    }
    if (bestscorerow > maxdist) goto done;
  }
  {
    int bestpos = (int)(patlen - (unsigned)kk);
    if (textlen < (unsigned)bestpos) bestpos = (int)textlen;
    if (bestpos < 0) bestpos = 0;
    int bestval = (int)dp[PIDX(patlen, bestpos)];
    ub = (int)patlen + kk;
    if (textlen < (unsigned)ub) ub = (int)textlen;
    for (int k2 = bestpos + 1; k2 <= ub; k2++) {
      const int val = (int)dp[PIDX(patlen, k2)];
      if (val < bestval || (val <= bestval && (best[PIDX(patlen, k2)] & PA_M))) { bestval = val; bestpos = k2; }
    }
    p = (int)patlen;
    t = bestpos;
    if (t < p - kk || t > p + kk) goto done;
    *matchlen = bestpos;
    *value = bestval;
    ret = 1;
  }
done:
  free(dp);
  free(best);
  return ret;
#undef PIDX
}

/* ------------------------------------------------------------------------- */
/* exact_halves::find_patterns, exact_halves.cc:118-199 (+ init, 201-224):    */
/* exact occurrences of the two halves of every pattern (inner index ids:     */
/* left half 2i-1, right half 2i), taken in the order (position, inner id     */
/* descending); a left-half occurrence ending at pos is extended to the right */
/* (primer_alignment_lmatch, primer_alignment.cc:581-658: text [pos, pos +    */
/* |right| + k), end = pos + matchlen), a right-half occurrence ending at pos */
/* to the left (primer_alignment_rmatch, 660-741: text [max(0, pos - m - k),  */
/* pos - |right|), end = pos); a hit is reported when its end lies beyond the */
/* pattern's last reported end (+ 2k with indels).  half_hits: occurrences of */
/* the 2*npat half patterns (pid = inner id), any order.  Output (end, pid,   */
/* val) with pid = outer pattern id.                                          */
static int half_hit_cmp(const void *a, const void *b) {
  const pmo_hit *x = (const pmo_hit *)a, *y = (const pmo_hit *)b;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return x->pid > y->pid ? -1 : (x->pid < y->pid ? 1 : 0);
}
long pmo_exact_halves(const uint8_t *T, long L, int npat, const char *const *pats,
                      pmo_hit *half_hits, long nhalf, int k, uint8_t eos, int wc, int tn, int indels,
                      const int *lconst, const int *rconst, pmo_hit *out, long cap) {
  qsort(half_hits, (size_t)nhalf, sizeof(pmo_hit), half_hit_cmp);
  int64_t *lasthit = (int64_t *)calloc((size_t)npat + 1, sizeof(int64_t));
  long n = 0;
  uint8_t buf[512];
  for (long i = 0; i < nhalf; i++) {
    const int hid = half_hits[i].pid;          /* inner id */
    const int pid = (hid + 1) / 2;             /* outer pattern */
    const int64_t pos = half_hits[i].pos;
    const char *pat = pats[pid - 1];
    const int m = (int)strlen(pat), l1 = m / 2, l2 = m - l1;
    const unsigned esb = (unsigned)(lconst ? lconst[pid - 1] : 0), eeb = (unsigned)(rconst ? rconst[pid - 1] : 0);
    int matchlen = 0, value = 0, ok;
    int64_t end;
    if (hid % 2 == 1) {
      const int buflen = l2 + k;
      for (int j = 0; j < buflen; j++) buf[j] = stream_at(T, L, pos + j);
      ok = pmo_primer_global_align(buf, (unsigned)buflen, pat + l1, 1, esb - (unsigned)l1, eeb, k, eos, wc, tn,
                                   indels, &matchlen, &value);
      end = pos + matchlen;
    } else {
      int64_t textstart = 0;
      const int patlen = l1 + l2 + k;
      if (pos > (int64_t)patlen) textstart = pos - patlen;
      const int buflen = (int)(pos - l2 - textstart);
      char left[80];
      memcpy(left, pat, (size_t)l1);
      left[l1] = 0;
      for (int j = 0; j < buflen; j++) buf[j] = stream_at(T, L, textstart + j);
      ok = pmo_primer_global_align(buf, (unsigned)(buflen < 0 ? 0 : buflen), left, -1, esb, eeb - (unsigned)l2, k, eos, wc,
                                   tn, indels, &matchlen, &value);
      end = pos;
    }
    if (ok && end > lasthit[pid] + (indels ? 2 * k : 0)) {
      if (n < cap) { out[n].pos = end; out[n].pid = pid; out[n].val = value; }
      n++;
      lasthit[pid] = end;
    }
  }
  free(lasthit);
  return n;
}

/* ------------------------------------------------------------------------- */
/* exact_bases::find_patterns, exact_bases.cc:68-124 (+ init, 126-159): what  */
/* pick_pattern_index takes when the exact-base constraints (-s/-e/-5/-3) are */
/* at least 6 long (select.cc:131-133).  The exact part of pattern i is its   */
/* first esb characters (esb >= eeb) or its last eeb characters; every exact  */
/* occurrence of it (part_hits, pid = pattern id) is extended over the rest   */
/* by primer_alignment_lmatch / _rmatch; every success is a hit (no dedupe).  */
long pmo_exact_bases(const uint8_t *T, long L, int npat, const char *const *pats,
                     const pmo_hit *part_hits, long nparts, int k, uint8_t eos, int wc, int tn, int indels,
                     const int *lconst, const int *rconst, pmo_hit *out, long cap) {
  (void)npat;
  long n = 0;
  uint8_t buf[512];
  char part[80];
  for (long i = 0; i < nparts; i++) {
    const int pid = part_hits[i].pid;
    const int64_t pos = part_hits[i].pos;
    const char *pat = pats[pid - 1];
    const int m = (int)strlen(pat), esb = lconst[pid - 1], eeb = rconst[pid - 1];
    int matchlen = 0, value = 0, ok;
    int64_t end;
    if (esb >= eeb) {
      const int l1 = esb, l2 = m - esb, buflen = l2 + k;
      for (int j = 0; j < buflen; j++) buf[j] = stream_at(T, L, pos + j);
      ok = pmo_primer_global_align(buf, (unsigned)buflen, pat + l1, 1, (unsigned)esb - (unsigned)l1, (unsigned)eeb, k,
                                   eos, wc, tn, indels, &matchlen, &value);
      end = pos + matchlen;
    } else {
      const int l2 = eeb, l1 = m - eeb;
      int64_t textstart = 0;
      const int patlen = l1 + l2 + k;
      if (pos > (int64_t)patlen) textstart = pos - patlen;
      const int buflen = (int)(pos - l2 - textstart);
      memcpy(part, pat, (size_t)l1);
      part[l1] = 0;
      for (int j = 0; j < buflen; j++) buf[j] = stream_at(T, L, textstart + j);
      ok = pmo_primer_global_align(buf, (unsigned)(buflen < 0 ? 0 : buflen), part, -1, (unsigned)esb,
                                   (unsigned)eeb - (unsigned)l2, k, eos, wc, tn, indels, &matchlen, &value);
      end = pos;
    }
    if (ok) {
      if (n < cap) { out[n].pos = end; out[n].pid = pid; out[n].val = value; }
      n++;
    }
  }
  return n;
}
