// pm_halves.cuh -- exact_halves and exact_bases semantics on the device (sm_100a).
//
// For one mismatch / edit the reference does not use filter_bitvec:
// pick_pattern_index (select.cc:120-126) wraps its exact index in exact_halves
// (exact_halves.cc:118-199).  With one error one HALF of the pattern occurs
// exactly; every exact occurrence of a half is extended over the other half by
// primer_alignment (primer_alignment_lmatch / _rmatch, primer_alignment.cc:581-741,
// DP = primer_alignment::global_align, lines 9-462, yes/no mode), and a hit is
// reported when its end lies beyond the pattern's previously reported end (+ 2k
// with indels).  Here:
//
//   1. the exact occurrences of the 2N halves come from an inner context (the
//      same pair-seed index scan as every k = 0 search);
//   2. halves_extend_kernel: one thread per occurrence runs the banded DP with
//      two rolling rows in local memory; the few that extend are appended to a
//      compact list of (end, value);
//   3. the records are sorted by (pattern, position, right half first) -- the
//      order exact_halves.cc:113-116 gives them -- and halves_report_kernel walks
//      each pattern's run with the reference's `lasthit` rule, one thread per
//      pattern (the rule is sequential per pattern, the runs are short);
//   4. reported (pattern, end) pairs go through the same final alignment as every
//      other index (primer_match.cc:1132-1153).
//
// exact_bases (exact_bases.cc, select.cc:131-133: the automatic choice when the -s/-e/-5/-3
// constraints are at least 6 characters) is the same machine with the constrained prefix or
// suffix of the pattern as the exact part and without step 3: every extension is a hit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pm_align.cuh"

namespace pm {

constexpr int HALF_MAX = 72;  // longest part + k + 1 (patterns <= 64, k <= 3)

struct HalfRec {
  int64_t end;      // pal.end() / par.end()
  int32_t value;    // DP value
  uint32_t ok;      // global_align's bool
};

// primer_alignment::global_align (primer_alignment.cc:9-462) with yesno_ = true and
// dna_mut_ = false.  text[0 .. textlen) against pat[0 .. patlen), from the left
// (dirn > 0) or from the right (dirn < 0).  lmatch / rmatch arrive as the reference's
// unsigned ints and land in int lbexact / rbexact exactly as there (lines 43-52).
// Line 97 reads `patch` before it is assigned; the comparison is false here, as the
// goldens made with the reference binary confirm.
__device__ __forceinline__ bool primer_global_align(const AlignEnv &env, const uint8_t *text, int64_t text_n,
                                                    int64_t text0, int textlen, const uint8_t *pat, int patlen,
                                                    int dirn, uint32_t lmatch, uint32_t rmatch, int &matchlen,
                                                    int &value) {
  const int maxdist = env.k;
  const int kk = env.indels ? maxdist : 0;
  const int cv = 5 * maxdist + 1;
  const uint32_t eos = env.eos;
  int lbexact = 0, rbexact = patlen + 1;
  if (dirn < 0) {
    if (lmatch > 0) rbexact = (int)((uint32_t)(patlen + 1) - lmatch);
    if (rmatch > 0) lbexact = (int)rmatch;
  } else {
    if (lmatch > 0) lbexact = (int)lmatch;
    if (rmatch > 0) rbexact = (int)((uint32_t)(patlen + 1) - rmatch);
  }
  auto tx = [&](int i) -> uint32_t {  // text[i]; characters past the stream read as 0
    const int64_t at = text0 + i;
    return (at >= 0 && at < text_n) ? (uint32_t)text[at] : 0u;
  };
  // two rolling rows of the band; column t of row p lives at [t]
  int prev[HALF_MAX + 4], cur[HALF_MAX + 4];
  uint64_t curM = 0;  // bit t: cell (p, t) can be reached by a diagonal (match-type) move
#pragma unroll 1
  for (int t = 0; t <= textlen && t < HALF_MAX + 4; t++) prev[t] = 0x3fffffff;
  prev[0] = 0;
  {
    int ub = kk;
    if (textlen < ub) ub = textlen;
    for (int t = 1; t <= ub; t++) {
      const uint32_t textch = dirn > 0 ? tx(t - 1) : tx(textlen - t);
      if (!env.indels || 0 < lbexact || 0 >= rbexact || textch == eos) prev[t] = cv;
      else prev[t] = prev[t - 1] + 1;
    }
  }
  int col0 = 0;  // dp[p][0] of the previous row
  for (int p = 1; p <= patlen; p++) {
    // dp[p][0] (lines 60-79): only rows 1 .. min(kk, patlen) are defined
    int c0 = 0x3fffffff;
    if (p <= kk) {
      if (!env.indels || p < lbexact || p >= rbexact) c0 = cv;
      else c0 = col0 + 1;
    }
    int lb = p - kk;
    if (lb < 1) lb = 1;
    int ub = p + kk;
    if (textlen < ub) ub = textlen;
    int bestscorerow = cv;
    curM = 0;
    const uint32_t patch = dirn > 0 ? pat[p - 1] : pat[patlen - p];
    for (int t = 0; t <= textlen && t < HALF_MAX + 4; t++) cur[t] = 0x3fffffff;
    cur[0] = c0;
    for (int t = lb; t <= ub; t++) {
      const uint32_t textch = dirn > 0 ? tx(t - 1) : tx(textlen - t);
      const int diag = (t - 1 == 0) ? col0 : prev[t - 1];
      int v, v1;
      bool m;
      if (textch == patch) { v = diag; m = true; }
      else if (env.wc && compat_bit(env.compat, textch, patch) && (env.tn || textch != 'N')) { v = diag; m = true; }
      else if (textch == eos || patch == eos || p <= lbexact || p >= rbexact) { v = cv; m = false; }
      else { v = diag + 1; m = true; }
      // insertion: dp[p][t-1] + 1
      if (textch == eos || patch == eos || !env.indels || t <= lb || p < lbexact || p >= rbexact) v1 = cv;
      else v1 = cur[t - 1] + 1;
      if (v1 < v) { v = v1; m = false; }
      // deletion: dp[p-1][t] + 1
      if (!env.indels || t >= ub || p <= lbexact || p >= rbexact) v1 = cv;
      else v1 = prev[t] + 1;
      if (v1 < v) { v = v1; m = false; }
      cur[t] = v;
      if (m) curM |= 1ull << (t & 63);
      if (bestscorerow > v) bestscorerow = v;
    }
    if (bestscorerow > maxdist) return false;
    col0 = c0;
    for (int t = 0; t <= textlen && t < HALF_MAX + 4; t++) prev[t] = cur[t];
  }
  int bestpos = patlen - kk;
  if (textlen < bestpos) bestpos = textlen;
  if (bestpos < 0) bestpos = 0;
  int bestval = (bestpos == 0) ? col0 : prev[bestpos];
  int ub = patlen + kk;
  if (textlen < ub) ub = textlen;
  for (int k2 = bestpos + 1; k2 <= ub; k2++) {
    const int val = prev[k2];
    if (val < bestval || (val <= bestval && ((curM >> (k2 & 63)) & 1ull))) { bestval = val; bestpos = k2; }
  }
  if (bestpos < patlen - kk || bestpos > patlen + kk) return false;
  matchlen = bestpos;
  value = bestval;
  return true;
}

// one thread per exact occurrence of an inner pattern.  keys[i] = inner slot << 40 | pos.
//   bases == 0 (exact_halves): inner_id[inner slot] = 2 * outer slot + 1 (left half) or + 2
//     (right half); the occurrences that extend are appended to rec / sortkey.
//   bases != 0 (exact_bases, exact_bases.cc:68-124): inner_id = outer slot + 1; the exact part
//     is the first esb characters (esb >= eeb) or the last eeb; every occurrence that extends
//     is a hit at once (no lasthit rule): appended to out when its end lies in [own_lo, own_hi).
__global__ void __launch_bounds__(128)
halves_extend_kernel(const AlignEnv env, const unsigned long long *__restrict__ keys,
                     const unsigned long long *__restrict__ n_ptr, uint64_t bound,
                     const uint32_t *__restrict__ inner_id, int bases, HalfRec *__restrict__ rec,
                     unsigned long long *__restrict__ sortkey, unsigned long long *__restrict__ n_ok,
                     int64_t own_lo, int64_t own_hi, RawHit *__restrict__ out, unsigned long long *__restrict__ n_out,
                     uint64_t cap) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound || i >= *n_ptr) return;   // the occurrence count stays on the device
  const unsigned long long key = keys[i];
  const uint32_t hid = inner_id[(uint32_t)(key >> 40)];
  const int64_t pos = (int64_t)(key & 0xffffffffffULL);
  const uint32_t slot = bases ? hid - 1 : (hid - 1) >> 1;   // outer slot
  const uint8_t *pat = env.pat_chars + (size_t)slot * PAT_STRIDE;
  const int m = env.plen[slot] & 0x7f;
  const int esb_i = (int)env.lconst[slot], eeb_i = (int)env.rconst[slot];
  const uint32_t esb = (uint32_t)esb_i, eeb = (uint32_t)eeb_i;
  // the LEFT part matched exactly: odd inner id (halves), the prefix is the exact part (bases)
  const bool left = bases ? (esb_i >= eeb_i) : ((hid & 1u) != 0);
  const int l1 = bases ? (left ? esb_i : m - eeb_i) : m / 2, l2 = m - l1;
  HalfRec r;
  int matchlen = 0, value = 0;
  if (left) {
    // primer_alignment_lmatch::align: text [pos, pos + |right| + k), right part from the left
    const int buflen = l2 + env.k;
    r.ok = primer_global_align(env, env.text, env.n, pos, buflen, pat + l1, l2, 1, esb - (uint32_t)l1, eeb,
                               matchlen, value) ? 1u : 0u;
    r.end = pos + matchlen;
  } else {
    // primer_alignment_rmatch::align: text [max(0, pos - m - k), pos - |right|), left part from the right
    int64_t textstart = 0;
    const int patlen = l1 + l2 + env.k;
    if (pos > (int64_t)patlen) textstart = pos - patlen;
    int buflen = (int)(pos - l2 - textstart);
    if (buflen < 0) buflen = 0;
    r.ok = primer_global_align(env, env.text, env.n, textstart, buflen, pat, l1, -1, esb, eeb - (uint32_t)l2,
                               matchlen, value) ? 1u : 0u;
    r.end = pos;
  }
  r.value = value;
  if (!r.ok) return;   // most occurrences do not extend: only the others are kept
  if (bases) {
    if (r.end >= own_lo && r.end < own_hi) {
      const unsigned long long at = atomicAdd(n_out, 1ULL);
      if (at < cap) {
        RawHit h;
        h.end = r.end;
        h.first = r.end;
        h.slot = slot;
        h.val = r.value;
        // exact_bases.cc:68-124: the occurrences of the exact parts in (position, pattern) order
        h.ord = ((unsigned long long)pos << 24) | (unsigned long long)((slot + 1u) & 0xffffffu);
        out[at] = h;
      }
    }
    return;
  }
  const unsigned long long at = atomicAdd(n_ok, 1ULL);
  rec[at] = r;
  // exact_halves.cc:113-116: by position, then inner id DESCENDING (the right half first)
  sortkey[at] = ((unsigned long long)slot << 42) | ((unsigned long long)pos << 1) | (left ? 1ull : 0ull);
}

// one thread per pattern: its run of sorted records, the `lasthit` rule of
// exact_halves.cc:147-171.  own_lo / own_hi: report the hits whose end lies in there.
__global__ void __launch_bounds__(128)
halves_report_kernel(const unsigned long long *__restrict__ sortkey, const uint32_t *__restrict__ order,
                     const HalfRec *__restrict__ rec, const unsigned long long *__restrict__ n_ptr,
                     uint32_t n_slots, int k, int indels,
                     int64_t own_lo, int64_t own_hi, RawHit *__restrict__ out, unsigned long long *__restrict__ n_out,
                     uint64_t cap) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const uint64_t n = *n_ptr;   // records that extended (the sorted list is padded behind them)
  // first record of this pattern: lower bound of slot << 42
  uint64_t lo = 0, hi = n;
  const unsigned long long want = (unsigned long long)slot << 42;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (sortkey[mid] < want) lo = mid + 1; else hi = mid;
  }
  int64_t lasthit = 0;
  const int64_t slack = indels ? 2 * k : 0;
  for (uint64_t i = lo; i < n && (sortkey[i] >> 42) == slot; i++) {
    const HalfRec r = rec[order[i]];
    if (!r.ok) continue;
    if (r.end > lasthit + slack) {
      lasthit = r.end;
      if (r.end >= own_lo && r.end < own_hi) {
        const unsigned long long at = atomicAdd(n_out, 1ULL);
        if (at < cap) {
          RawHit h;
          h.end = r.end;
          h.first = r.end;
          h.slot = slot;
          h.val = r.value;
          // exact_halves.cc:113-116: half occurrences by position, then inner id DESCENDING
          // (inner id = 2 * slot + 1 for the left half, + 2 for the right half)
          const unsigned long long sk = sortkey[i];
          const unsigned long long hid = 2ull * slot + ((sk & 1ull) ? 1ull : 2ull);
          h.ord = (((sk >> 1) & 0xffffffffffULL) << 24) | (0xffffffull - (hid & 0xffffffull));
          out[at] = h;
        }
      }
    }
  }
}

}  // namespace pm
