// pm_align.cuh -- candidate chains and the banded edit-distance alignment
// (sm_100a; one thread per chain / hit, the populations are tiny next to the
// dense scan).
//
// filter_bitvec::find_patterns (filter_bitvec.cc:126-236) groups consecutive
// candidates of one pattern that are at most 2k+1 apart into a chain and asks
// editdist_alignment (yes/no form) whether the pattern aligns with <= k errors
// ending somewhere in [first, last].  primer_match then re-aligns every
// reported hit at (end, end) to get the numbers it prints
// (primer_match.cc:1132-1153).  Both alignments are the same dynamic programme,
// editdist_alignment::align (pattern_alignment.cc:117-705, dna_mut=false,
// trans=false): pattern and text are walked right-to-left, row p = p-th
// pattern character from the right, column t = t-th text character from the
// right end of the buffer, cells outside the band |t-p| <= k (+ the free end
// range) are never touched, ties are remembered as a bit set per cell and
// resolved by the traceback's preference rules.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PM_HD __host__ __device__
#else
#define PM_HD
#endif

namespace pm {

// alignment_code.h:5-19
enum : uint8_t {
  AC_NONE = 0, AC_EQUAL = 1, AC_WC_EQUAL = 2, AC_SUBST = 3, AC_INSERT = 4,
  AC_DELETE = 5, AC_CONSTRAINT = 6, AC_END = 7
};
// tie sets of the traceback matrix
enum : uint8_t {
  M_EQUAL = 2, M_WC = 4, M_SUBST = 8, M_INSERT = 16, M_DELETE = 32, M_CONSTRAINT = 64, M_END = 128
};

struct AlignEnv {
  const uint8_t *text;       // shard bytes
  int64_t n;                 // shard length
  const uint8_t *pat_chars;  // [n_slots][PAT_STRIDE]
  const uint8_t *plen;       // [n_slots]
  const int16_t *lconst;     // [n_slots] exact_start_bases
  const int16_t *rconst;     // [n_slots] exact_end_bases
  const uint32_t *compat;    // 128x128 bit matrix: iupac_compatible(row, col), util.cc:121-185
  int k, indels, wc, tn;
  uint32_t eos;
};
constexpr int PAT_STRIDE = 64;

struct AlignResult {
  int64_t start, end;
  int32_t val, editdist;  // editdist -1 == MAXINT (constraint violation)
  int32_t ok;             // the function's bool
  int32_t ncodes;
};

PM_HD inline bool compat_bit(const uint32_t *compat, uint32_t row, uint32_t col) {
  if (row >= 128 || col >= 128) return false;
  return (compat[row * 4 + (col >> 5)] >> (col & 31)) & 1u;
}

// workspace bytes one alignment needs: dp as uint16 + tie sets as uint8
PM_HD inline uint64_t align_ws_bytes(int m, int64_t buflen) {
  return (uint64_t)(m + 1) * (uint64_t)(buflen + 1) * 3ull;
}
PM_HD inline int64_t align_textstart(int64_t e1, int m, int k) {
  return (e1 > (int64_t)m + k) ? e1 - m - k : 0;  // pattern_alignment.cc:137-139
}

// The plain statement of the dynamic programme over the full (m+1) x (buflen+1) matrix.  The
// kernels run the banded form below; this one stays as what tools/align_band_check.cc compares it
// with.  codes (if not null) receives the alignment codes in pattern order.
PM_HD inline void editdist_align(const AlignEnv &env, uint32_t slot, int64_t e1, int64_t e2,
                                      uint8_t *ws, uint8_t *codes, int max_codes, AlignResult &r) {
  const int m = env.plen[slot] & 0x7f;
  const uint8_t *pat = env.pat_chars + (size_t)slot * PAT_STRIDE;
  const int k = env.k;
  const int cvp = 5 * k + 1;  // pattern_alignment.cc:131
  const int64_t textstart = align_textstart(e1, m, k);
  const int64_t buflen = e2 - textstart;  // line 148
  const int kk = env.indels ? k : 0;
  const int64_t R = e2 - e1;
  r.ok = 0; r.editdist = -1; r.ncodes = 0; r.start = 0; r.end = 0; r.val = 0;
  if (buflen < 0) return;
  const int64_t W = buflen + 1;
  uint16_t *dp = reinterpret_cast<uint16_t *>(ws);
  uint8_t *best = ws + (uint64_t)(m + 1) * (uint64_t)W * 2ull;
  const int lconst = env.lconst[slot], rconst = env.rconst[slot];
  int lbexact = 0, rbexact = m + 1;  // lines 230-233
  if (lconst > 0) rbexact = m + 1 - lconst;
  if (rconst > 0) lbexact = rconst;
  const uint32_t eos = env.eos;
#define PM_IDX(p, t) ((int64_t)(p) * W + (int64_t)(t))
#define PM_BUF(t) ((textstart + buflen - (t)) >= 0 && (textstart + buflen - (t)) < env.n \
                       ? (uint32_t)env.text[textstart + buflen - (t)] : 0u)
#define PM_PAT(p) ((uint32_t)pat[m - (p)])
  dp[PM_IDX(0, 0)] = 0; best[PM_IDX(0, 0)] = M_END;  // lines 242-243
  int64_t ub = kk; if (ub > m) ub = m;
  for (int64_t p = 1; p <= ub; p++) {  // lines 253-268
    if (!env.indels || p < lbexact || p >= rbexact || PM_PAT(p) == eos) {
      dp[PM_IDX(p, 0)] = (uint16_t)cvp; best[PM_IDX(p, 0)] = M_CONSTRAINT;
    } else {
      dp[PM_IDX(p, 0)] = dp[PM_IDX(p - 1, 0)] + 1; best[PM_IDX(p, 0)] = M_DELETE;
    }
  }
  ub = R + kk; if (buflen < ub) ub = buflen;
  for (int64_t t = 1; t <= ub; t++) {  // lines 276-294
    if (t <= R) {
      dp[PM_IDX(0, t)] = 0; best[PM_IDX(0, t)] = M_END;
    } else if (!env.indels || lbexact > 0) {
      dp[PM_IDX(0, t)] = (uint16_t)cvp; best[PM_IDX(0, t)] = M_CONSTRAINT;
    } else {
      dp[PM_IDX(0, t)] = dp[PM_IDX(0, t - 1)] + 1; best[PM_IDX(0, t)] = M_INSERT;
    }
  }
  bool bail = false;
  for (int64_t p = 1; p <= m && !bail; p++) {  // lines 296-437
    int64_t lb = p - kk; if (lb < 1) lb = 1;
    ub = p + R + kk; if (buflen < ub) ub = buflen;
    int bestrow = cvp;
    const uint32_t pc = PM_PAT(p);
    for (int64_t t = lb; t <= ub; t++) {
      int v, v1; uint8_t ac;
      const uint32_t tc = PM_BUF(t);
      if (tc == pc) {
        v = dp[PM_IDX(p - 1, t - 1)]; ac = M_EQUAL;
      } else if (env.wc && compat_bit(env.compat, pc, tc) && (tc != 'N' || env.tn)) {
        v = dp[PM_IDX(p - 1, t - 1)]; ac = M_WC;
      } else if (tc == eos || pc == eos || p <= lbexact || p >= rbexact) {
        v = cvp; ac = M_CONSTRAINT;
      } else {
        v = dp[PM_IDX(p - 1, t - 1)] + 1; ac = M_SUBST;
      }
      // text character without a pattern character, lines 366-391
      if (tc == eos || pc == eos || !env.indels || t <= lb || p < lbexact || p >= rbexact) {
        v1 = cvp;
        if (v1 < v) { v = v1; ac = M_CONSTRAINT; }
      } else {
        v1 = dp[PM_IDX(p, t - 1)] + 1;
        if (v1 < v) { v = v1; ac = M_INSERT; } else if (v1 == v) ac |= M_INSERT;
      }
      // pattern character without a text character, lines 392-416
      if (!env.indels || pc == eos || t >= ub || p <= lbexact || p >= rbexact) {
        v1 = cvp;
        if (v1 < v) { v = v1; ac = M_CONSTRAINT; }
      } else {
        v1 = dp[PM_IDX(p - 1, t)] + 1;
        if (v1 < v) { v = v1; ac = M_DELETE; } else if (v1 == v) ac |= M_DELETE;
      }
      dp[PM_IDX(p, t)] = (uint16_t)v; best[PM_IDX(p, t)] = ac;
      if (v < bestrow) bestrow = v;
    }
    if (bestrow > k) bail = true;  // lines 425-436
  }
  if (bail) {
    if (codes && max_codes > 0) codes[0] = AC_CONSTRAINT;
    r.ncodes = 1;
    return;
  }
  int64_t beststart = m - kk;  // lines 443-449
  if (beststart > buflen) beststart = buflen;
  if (beststart < 0) beststart = 0;
  int bestval = dp[PM_IDX(m, beststart)];
  ub = m + R + kk; if (buflen < ub) ub = buflen;
  for (int64_t x = beststart + 1; x <= ub; x++) {  // lines 461-475
    const int d = dp[PM_IDX(m, x)];
    if (d < bestval || (d <= bestval && (best[PM_IDX(m, x)] & (M_EQUAL | M_WC | M_SUBST)))) {
      bestval = d;
      beststart = x;
    }
  }
  int64_t p = m, t = beststart;
  if (t < p - kk || t > p + kk + R) {  // lines 482-490
    if (codes && max_codes > 0) codes[0] = AC_CONSTRAINT;
    r.ncodes = 1;
    return;
  }
  int lastac = AC_NONE;  // lines 509-590
  int n_sub = 0, n_ins = 0, n_del = 0, n_con = 0, nc = 0;
  while (!(best[PM_IDX(p, t)] & M_END)) {
    const int ac = best[PM_IDX(p, t)];
    const int match = ac & (M_EQUAL | M_WC | M_SUBST);
    const int wcb = ac & M_WC, sub = ac & M_SUBST, ins = ac & M_INSERT, del = ac & M_DELETE;
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
    else { r.ok = -1; return; }  // assert(0) in the reference
    if (lastac == AC_SUBST) n_sub++;
    else if (lastac == AC_INSERT) n_ins++;
    else if (lastac == AC_DELETE) n_del++;
    else if (lastac == AC_CONSTRAINT) n_con++;
    if (codes && nc < max_codes) codes[nc] = (uint8_t)lastac;
    nc++;
  }
  r.ncodes = nc;
  r.start = e2 - beststart;              // line 608
  r.end = r.start + (beststart - t);     // lines 603-609
  r.val = bestval;
  r.editdist = (n_con > 0) ? -1 : (n_sub + n_ins + n_del);  // pattern_alignment.h:108-123
  r.ok = (bestval <= k) ? 1 : 0;
#undef PM_IDX
#undef PM_BUF
#undef PM_PAT
}

// ---------------------------------------------------------------------------
// The same dynamic programme with BANDED storage.  Row p only ever touches the columns
// t in [p - kk, p + R + kk] (R = e2 - e1, kk = k with indels, else 0), so a cell lives at
// band index j = t - (p - kk), 0 <= j < Wb = R + 2 kk + 1:
//   dp    two rolling rows (only rows p-1 and p are read): dp[p-1][t-1] -> prev[j],
//         dp[p-1][t] -> prev[j+1], dp[p][t-1] -> cur[j-1];
//   best  the tie sets of the traceback, (m+1) x Wb bytes.
// A chain of a few candidates needs Wb <= 4k+1 + a few: the tie sets of such an alignment fit
// a per-thread array (ALIGN_BAND_MAX) and no global workspace is needed at all; long chains
// (repeats) take their storage from a pool.  Same cells, same order, same tie rules as
// editdist_align above -- tools/align_band_check.cc compares the two on random inputs on the
// CPU, the GPU tests compare the results with the reference's records.
constexpr int ALIGN_BAND_MAX = 20;
constexpr int ALIGN_LOCAL_BEST = (64 + 1) * ALIGN_BAND_MAX;   // patterns up to 64 characters

PM_HD inline int64_t align_band_width(int64_t e1, int64_t e2, int k, int indels) {
  return (e2 - e1) + 2 * (int64_t)(indels ? k : 0) + 1;
}
// pool bytes of one banded alignment: tie sets + two dp rows (int)
PM_HD inline uint64_t align_band_bytes(int m, int64_t wb) {
  return ((uint64_t)(m + 1) * (uint64_t)wb + 2ull * (uint64_t)(wb + 2) * 4ull + 31ull) & ~15ull;
}

PM_HD inline void editdist_align_band(const AlignEnv &env, uint32_t slot, int64_t e1, int64_t e2, uint8_t *best,
                                      int *row_a, int *row_b, uint8_t *codes, int max_codes, AlignResult &r) {
  const int m = env.plen[slot] & 0x7f;
  const uint8_t *pat = env.pat_chars + (size_t)slot * PAT_STRIDE;
  const int k = env.k;
  const int cvp = 5 * k + 1;
  const int64_t textstart = align_textstart(e1, m, k);
  const int64_t buflen = e2 - textstart;
  const int kk = env.indels ? k : 0;
  const int64_t R = e2 - e1;
  const int64_t Wb = R + 2 * kk + 1;
  r.ok = 0; r.editdist = -1; r.ncodes = 0; r.start = 0; r.end = 0; r.val = 0;
  if (buflen < 0) return;
  const int lconst = env.lconst[slot], rconst = env.rconst[slot];
  int lbexact = 0, rbexact = m + 1;
  if (lconst > 0) rbexact = m + 1 - lconst;
  if (rconst > 0) lbexact = rconst;
  const uint32_t eos = env.eos;
#define PM_B(p, t) best[(int64_t)(p) * Wb + ((int64_t)(t) - ((int64_t)(p) - kk))]
#define PM_BUF(t) ((textstart + buflen - (t)) >= 0 && (textstart + buflen - (t)) < env.n \
                       ? (uint32_t)env.text[textstart + buflen - (t)] : 0u)
#define PM_PAT(p) ((uint32_t)pat[m - (p)])
  int *prev = row_a, *cur = row_b;   // prev[j]: row p-1 at band index j (t = j + p-1 - kk)
  // row 0: t = j - kk
  prev[kk] = 0; PM_B(0, 0) = M_END;
  int64_t ub = R + kk; if (buflen < ub) ub = buflen;
  for (int64_t t = 1; t <= ub; t++) {
    if (t <= R) {
      prev[t + kk] = 0; PM_B(0, t) = M_END;
    } else if (!env.indels || lbexact > 0) {
      prev[t + kk] = cvp; PM_B(0, t) = M_CONSTRAINT;
    } else {
      prev[t + kk] = prev[t + kk - 1] + 1; PM_B(0, t) = M_INSERT;
    }
  }
  const int64_t col0_rows = (kk < m) ? kk : m;
  bool bail = false;
  for (int64_t p = 1; p <= m && !bail; p++) {
    // cur[j]: row p at t = j + p - kk
    if (p <= col0_rows) {   // column 0 of this row
      const int64_t j0 = kk - p;
      if (!env.indels || p < lbexact || p >= rbexact || PM_PAT(p) == eos) {
        cur[j0] = cvp; PM_B(p, 0) = M_CONSTRAINT;
      } else {
        cur[j0] = prev[j0 + 1] + 1; PM_B(p, 0) = M_DELETE;
      }
    }
    int64_t lb = p - kk; if (lb < 1) lb = 1;
    ub = p + R + kk; if (buflen < ub) ub = buflen;
    int bestrow = cvp;
    const uint32_t pc = PM_PAT(p);
    for (int64_t t = lb; t <= ub; t++) {
      const int64_t j = t - (p - kk);
      int v, v1; uint8_t ac;
      const uint32_t tc = PM_BUF(t);
      if (tc == pc) {
        v = prev[j]; ac = M_EQUAL;
      } else if (env.wc && compat_bit(env.compat, pc, tc) && (tc != 'N' || env.tn)) {
        v = prev[j]; ac = M_WC;
      } else if (tc == eos || pc == eos || p <= lbexact || p >= rbexact) {
        v = cvp; ac = M_CONSTRAINT;
      } else {
        v = prev[j] + 1; ac = M_SUBST;
      }
      if (tc == eos || pc == eos || !env.indels || t <= lb || p < lbexact || p >= rbexact) {
        v1 = cvp;
        if (v1 < v) { v = v1; ac = M_CONSTRAINT; }
      } else {
        v1 = cur[j - 1] + 1;
        if (v1 < v) { v = v1; ac = M_INSERT; } else if (v1 == v) ac |= M_INSERT;
      }
      if (!env.indels || pc == eos || t >= ub || p <= lbexact || p >= rbexact) {
        v1 = cvp;
        if (v1 < v) { v = v1; ac = M_CONSTRAINT; }
      } else {
        v1 = prev[j + 1] + 1;
        if (v1 < v) { v = v1; ac = M_DELETE; } else if (v1 == v) ac |= M_DELETE;
      }
      cur[j] = v; PM_B(p, t) = ac;
      if (v < bestrow) bestrow = v;
    }
    if (bestrow > k) bail = true;
    int *sw = prev; prev = cur; cur = sw;
  }
  if (bail) {
    if (codes && max_codes > 0) codes[0] = AC_CONSTRAINT;
    r.ncodes = 1;
    return;
  }
  // prev = row m, t = j + m - kk
  int64_t beststart = m - kk;
  if (beststart > buflen) beststart = buflen;
  if (beststart < 0) beststart = 0;
  int bestval = prev[beststart - (m - kk)];
  ub = m + R + kk; if (buflen < ub) ub = buflen;
  for (int64_t x = beststart + 1; x <= ub; x++) {
    const int d = prev[x - (m - kk)];
    if (d < bestval || (d <= bestval && (PM_B(m, x) & (M_EQUAL | M_WC | M_SUBST)))) {
      bestval = d;
      beststart = x;
    }
  }
  int64_t p = m, t = beststart;
  if (t < p - kk || t > p + kk + R) {
    if (codes && max_codes > 0) codes[0] = AC_CONSTRAINT;
    r.ncodes = 1;
    return;
  }
  int lastac = AC_NONE;
  int n_sub = 0, n_ins = 0, n_del = 0, n_con = 0, nc = 0;
  while (!(PM_B(p, t) & M_END)) {
    const int ac = PM_B(p, t);
    const int match = ac & (M_EQUAL | M_WC | M_SUBST);
    const int wcb = ac & M_WC, sub = ac & M_SUBST, ins = ac & M_INSERT, del = ac & M_DELETE;
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
    else { r.ok = -1; return; }
    if (lastac == AC_SUBST) n_sub++;
    else if (lastac == AC_INSERT) n_ins++;
    else if (lastac == AC_DELETE) n_del++;
    else if (lastac == AC_CONSTRAINT) n_con++;
    if (codes && nc < max_codes) codes[nc] = (uint8_t)lastac;
    nc++;
  }
  r.ncodes = nc;
  r.start = e2 - beststart;
  r.end = r.start + (beststart - t);
  r.val = bestval;
  r.editdist = (n_con > 0) ? -1 : (n_sub + n_ins + n_del);
  r.ok = (bestval <= k) ? 1 : 0;
#undef PM_B
#undef PM_BUF
#undef PM_PAT
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// Chains of candidates: keys are sorted (slot << 40 | pos) and unique; a chain continues while the
// next candidate of the pattern is at most 2k+1 past the previous one (filter_bitvec.cc:157-168).
struct Chain {
  int64_t first, last;  // candidate positions (local to the shard)
  uint32_t slot;
  uint32_t pad;
};

struct RawHit {
  int64_t end;       // pa.end() of the chain alignment
  int64_t first;     // chain head (ordering / ownership)
  uint32_t slot;
  int32_t val;
  uint64_t ord;      // exact_halves / exact_bases: the order the reference reports the hit in (position of
                     // the exact part's occurrence, then its inner id); unused for chains
};

// ---------------------------------------------------------------------------
// The tail with every count left on the device: the host only knows an upper bound
// (`bound` = candidates before duplicates were removed) and sizes buffers and grids by it.
struct TailCounts {
  const unsigned long long *n_cands;   // sorted unique candidates
  unsigned long long *n_chains;
  unsigned long long *n_raw;           // hits after the chain alignment
  unsigned long long *n_bogus;
  unsigned long long *cut;             // an owned chain reaches the end of the right guard
  unsigned long long *pool_used;       // bytes of pool taken (keeps counting past the capacity)
};

__global__ void chain_heads2_kernel(const unsigned long long *keys, const unsigned long long *n_ptr, uint64_t bound,
                                    int gap, uint32_t *head) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound) return;
  uint32_t h = 0;
  if (i < *n_ptr) {
    h = 1;
    if (i > 0) {
      const unsigned long long a = keys[i - 1], b = keys[i];
      if ((a >> 40) == (b >> 40) && (b & 0xffffffffffULL) - (a & 0xffffffffffULL) <= (unsigned long long)gap)
        h = 0;
    }
  }
  head[i] = h;
}

__global__ void chain_fill2_kernel(const unsigned long long *keys, const unsigned long long *n_ptr,
                                   const uint32_t *head, const uint32_t *chain_incl, Chain *chains,
                                   unsigned long long *n_chains) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n = *n_ptr;
  if (i >= n) return;
  const uint32_t c = chain_incl[i] - 1;
  const unsigned long long key = keys[i];
  if (head[i]) {
    chains[c].first = (int64_t)(key & 0xffffffffffULL);
    chains[c].slot = (uint32_t)(key >> 40);
    chains[c].pad = 0;
  }
  if (i + 1 == n || head[i + 1]) chains[c].last = (int64_t)(key & 0xffffffffffULL);
  if (i + 1 == n) *n_chains = (unsigned long long)c + 1ULL;
}

// one thread per chain: yes/no alignment over [first, last] (filter_bitvec.cc:170-236).  Chains
// whose band fits ALIGN_BAND_MAX keep everything in per-thread memory; the others take
// align_band_bytes() from the pool (bump allocation; a pool that is too small shows as
// pool_used > pool_cap on the host, which grows it and runs this kernel again).
__global__ void __launch_bounds__(64)
chain_align2_kernel(const AlignEnv env, const Chain *chains, TailCounts tc, uint64_t bound, int64_t own_lo,
                    int64_t own_hi, int64_t guard_end, uint8_t *pool, unsigned long long pool_cap, RawHit *hits) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound || i >= *tc.n_chains) return;
  const Chain c = chains[i];
  if (c.first < own_lo || c.first >= own_hi) return;   // another shard owns it
  if (guard_end > 0 && c.last + (2 * env.k + 1) >= guard_end) *tc.cut = 1ULL;
  const int m = env.plen[c.slot] & 0x7f;
  const int64_t wb = align_band_width(c.first, c.last, env.k, env.indels);
  AlignResult r;
  if (wb <= ALIGN_BAND_MAX) {
    uint8_t best[ALIGN_LOCAL_BEST];
    int row_a[ALIGN_BAND_MAX + 2], row_b[ALIGN_BAND_MAX + 2];
    editdist_align_band(env, c.slot, c.first, c.last, best, row_a, row_b, nullptr, 0, r);
  } else {
    const unsigned long long need = align_band_bytes(m, wb);
    const unsigned long long off = atomicAdd(tc.pool_used, need);
    if (off + need > pool_cap) return;   // the host sees pool_used > pool_cap
    uint8_t *ws = pool + off;
    int *rows = reinterpret_cast<int *>(ws + (((uint64_t)(m + 1) * (uint64_t)wb + 15ull) & ~15ull));
    editdist_align_band(env, c.slot, c.first, c.last, ws, rows, rows + (wb + 2), nullptr, 0, r);
  }
  if (r.ok == 1) {
    const unsigned long long at = atomicAdd(tc.n_raw, 1ULL);
    RawHit h;
    h.end = r.end; h.first = c.first; h.slot = c.slot; h.val = r.val; h.ord = 0;
    hits[at] = h;
  }
}
#endif  // __CUDACC__

}  // namespace pm
