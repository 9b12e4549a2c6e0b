// pm_lv.cuh -- seed verification of the edit scan by diagonal extension (sm_100a; the
// functions are host/device so that tools/lv_check.cc can run them against the oracle's
// restatement of the reference recurrence on the CPU).
//
// What has to be reproduced: the candidate set of shift_and_inexact::find_patterns with
// indels (shift_and_inexact.cc:253-326): every text position t at which level k of the
// Wu-Manber state holds the last pattern bit.  Read as a dynamic programme, D[i][c] = fewest
// errors with which the pattern prefix of length i ends at text column c:
//   D[0][c] = 0 (an occurrence may start anywhere); match / substitution / insertion /
//   deletion at unit cost; an end-of-sequence character cannot be consumed, and it clears
//   every level, so in the column right behind one D[i][.] = infinity for i >= 1: an
//   occurrence that starts at the first character of a sequence cannot begin by deleting
//   pattern characters (the reference's state after an end of sequence is all zero, not the
//   0..l-1 bits of its initial state);
//   a text character that matches no pattern symbol (N) can be substituted or inserted.
// Candidates: columns c with D[m][c] <= k.
//
// A seed of the pair index says: piece A (pattern offset a, length L) occurs exactly at text
// position p.  Every occurrence with <= k errors keeps two of its k+2 pieces intact, so the
// union over all seeds of "alignments that run through A's diagonal cells" is the candidate
// set.  Such an alignment splits at the anchor:
//   left   pattern[0, a) against the text ending at p, start free      -> d_l = fewest errors
//   right  pattern[a+L, m) against the text from p+L, end free         -> fewest errors per end
// and both halves are diagonal-wise furthest-reaching computations (Landau-Vishkin): at most
// (k+1)^2 extensions, each a count of trailing / leading zero bits of a mismatch vector of
// the packed text against the pattern's per-code accept masks.  Cost per seed: 2k+1 mismatch
// vectors + the extensions, instead of (m + 3k + 1) x (k+1) recurrence steps.
#pragma once
#include <stdint.h>
#include "pm_align.cuh"

namespace pm {

#if defined(__CUDA_ARCH__)
#define PM_FSR(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#define PM_CTZ64(x) (__ffsll((long long)(x)) - 1)
#define PM_CLZ64(x) __clzll((long long)(x))
#else
static inline uint32_t pm_fsr_host(uint32_t lo, uint32_t hi, uint32_t s) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31u));
}
#define PM_FSR(lo, hi, s) pm::pm_fsr_host((lo), (hi), (s))
#define PM_CTZ64(x) __builtin_ctzll((unsigned long long)(x))
#define PM_CLZ64(x) __builtin_clzll((unsigned long long)(x))
#endif

constexpr int LV_NONE = -64;   // "diagonal not reached"

// codes and flags of the 48 characters starting at text position s_nom - K (three packed words,
// character 0 at bit 0), and the pattern's accept masks
struct LvCtx {
  uint32_t c[3], f[3];
  uint64_t acc[4];     // bit 2j: pattern character j accepts text code 0 / 1 / 2 / 3
  uint64_t lenmask;    // bits below 2m
  uint64_t pcodes;     // plain patterns (every character accepts exactly one code): the 2-bit codes
  int plain;           // ... and the flag; lv_mis then is an XOR instead of four masked selects
  int m;
};

// mismatch vector of the diagonal fd (text position = s_nom + j + fd for pattern position j):
// bit 2j set <=> pattern character j does not match there (flagged text never matches)
PM_HD inline uint64_t lv_mis(const LvCtx &t, int fd, int K) {
  const uint32_t sh = 2u * (uint32_t)(fd + K);
  const uint64_t codes = (uint64_t)PM_FSR(t.c[0], t.c[1], sh) | ((uint64_t)PM_FSR(t.c[1], t.c[2], sh) << 32);
  const uint64_t flags = (uint64_t)PM_FSR(t.f[0], t.f[1], sh) | ((uint64_t)PM_FSR(t.f[1], t.f[2], sh) << 32);
  constexpr uint64_t M5 = 0x5555555555555555ULL;
  if (t.plain) {
    const uint64_t x = codes ^ t.pcodes;
    return (((x | (x >> 1)) | flags) & M5 & t.lenmask) | ~t.lenmask;
  }
  const uint64_t lo = codes & M5, hi = (codes >> 1) & M5;
  const uint64_t match = (~(hi | lo) & t.acc[0]) | (lo & ~hi & t.acc[1]) | (hi & ~lo & t.acc[2]) | (hi & lo & t.acc[3]);
  return ((~match | flags) & M5 & t.lenmask) | ~t.lenmask;   // positions past the pattern stop every extension
}

// matches from pattern position j upwards / from j downwards along a diagonal
PM_HD inline int lv_lce_fwd(uint64_t mis, int j) {
  if (j >= 32) return 0;
  const uint64_t v = mis >> (2 * j);
  return v ? (PM_CTZ64(v) >> 1) : (32 - j);
}
PM_HD inline int lv_lce_bwd(uint64_t mis, int j) {   // pattern positions j, j-1, ..
  if (j < 0) return 0;
  const uint64_t v = (mis & 0x5555555555555555ULL) << (62 - 2 * j);
  return v ? (PM_CLZ64(v) >> 1) : (j + 1);
}

// Ends of the occurrences of the pattern that run through the anchor (pattern offset a,
// length L, at text position s_nom + a): bit (d + K) of the result <=> an occurrence with
// <= K errors ends behind text position s_nom + m + d - 1, i.e. candidate position
// s_nom + m + d.  K is a template parameter so that the diagonal arrays stay in registers.
PM_HD inline void lv_set_pattern(LvCtx &t, int m, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3, int plain) {
  t.m = m;
  t.lenmask = (m >= 32) ? ~0ULL : ((1ULL << (2 * m)) - 1ULL);
  t.acc[0] = a0; t.acc[1] = a1; t.acc[2] = a2; t.acc[3] = a3;
  t.plain = plain;
  t.pcodes = (a1 | a3) | ((a2 | a3) << 1);
}

// The q-gram filter alone: false <=> no occurrence with <= K errors can run through this window
// (what lv_verify decides first).  The verify kernel asks it for every seed and runs lv_verify,
// whose cost a random seed never reaches, only for the few that pass, 32 of them at a time.
template <int K>
PM_HD inline bool lv_prefilter(const LvCtx &t) {
  constexpr int ND = 2 * K + 1;
  uint64_t any3 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < ND; i++) {
    const uint64_t nm = ~lv_mis(t, i - K, K);
    any3 |= nm & (nm >> 2) & (nm >> 4);
  }
#if defined(__CUDA_ARCH__)
  const int n3 = __popcll(any3 & 0x5555555555555555ULL);
#else
  const int n3 = __builtin_popcountll(any3 & 0x5555555555555555ULL);
#endif
  return n3 >= t.m - 2 - 3 * K;
}

template <int K>
PM_HD inline uint32_t lv_verify(const LvCtx &t, int a, int L) {
  const int m = t.m;
  constexpr int ND = 2 * K + 1;
  // the mismatch vectors of the 2K+1 diagonals (every loop below has compile-time bounds, so
  // the vectors and the furthest-reaching rows stay in registers)
  uint64_t mis[ND];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < ND; i++) mis[i] = lv_mis(t, i - K, K);
  // The seed's claim itself: the index keys on the 2-bit codes alone, and a flagged character (one
  // that matches nothing, an end of sequence) carries an arbitrary code that can complete a piece
  // by accident.  An anchor that is not an exact match is no anchor; a real occurrence has two
  // pieces that are, and is seeded by those.
  if ((mis[K] >> (2 * a)) & ((1ULL << (2 * L)) - 1ULL)) return 0u;
  // q-gram filter (q = 3): K errors cut the pattern into at most K+1 runs of matches, each on one
  // of these diagonals, and destroy at most 3K of its m-2 triples of consecutive characters; the
  // others match as a whole on some diagonal.  Random seeds fail here, before any extension.
  {
    uint64_t any3 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < ND; i++) {
      const uint64_t nm = ~mis[i];
      any3 |= nm & (nm >> 2) & (nm >> 4);
    }
#if defined(__CUDA_ARCH__)
    const int n3 = __popcll(any3 & 0x5555555555555555ULL);
#else
    const int n3 = __builtin_popcountll(any3 & 0x5555555555555555ULL);
#endif
    if (n3 < m - 2 - 3 * K) return 0u;
  }
  // walls: end-of-sequence characters of the window (flag bit 2i+1 of character i, window
  // offset i <-> text position s_nom - K + i)
  constexpr uint32_t ODD = 0xAAAAAAAAu;
  const uint32_t e0 = t.f[0] & ODD, e1 = t.f[1] & ODD, e2 = t.f[2] & ODD;
  // right wall wr (relative to s_nom): first end of sequence at or behind the anchor's end
  int wr = 1 << 20;
  {
    const int from = a + L + K;   // window offset
    for (int w = 0; w < 3; w++) {
      uint32_t e = (w == 0 ? e0 : (w == 1 ? e1 : e2));
      const int base = 16 * w;
      if (from > base) e = (from - base >= 16) ? 0u : (e & (0xffffffffu << (2 * (from - base))));
      if (e) {
#if defined(__CUDA_ARCH__)
        const int bit = __ffs((int)e) - 1;
#else
        const int bit = __builtin_ctz(e);
#endif
        wr = base + (bit >> 1) - K;
        break;
      }
    }
  }
  // left wall: last end of sequence before the anchor -> usable characters to its left
  int lw = 1 << 20;
  {
    const int upto = a + K;       // window offsets < upto
    for (int w = 2; w >= 0; w--) {
      uint32_t e = (w == 0 ? e0 : (w == 1 ? e1 : e2));
      const int base = 16 * w;
      if (upto <= base) continue;
      if (upto - base < 16) e &= ~(0xffffffffu << (2 * (upto - base)));
      if (e) {
#if defined(__CUDA_ARCH__)
        const int bit = 31 - __clz((int)e);
#else
        const int bit = 31 - __builtin_clz(e);
#endif
        lw = a - (base + (bit >> 1) - K + 1);
        break;
      }
    }
  }
  // ---- left: pattern[0, a) reversed against the text going left from the anchor ----------
  // state: j' pattern characters and x' = j' + d' text characters consumed; reversed diagonal
  // d' is the forward diagonal -d'.  x' <= lw; no deletion at x' == lw (see the header).
  int dl = 0;
  if (a > 0) {
    int F[ND], G[ND];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < ND; i++) F[i] = LV_NONE;
    {
      int j = lv_lce_bwd(mis[K], a - 1);
      if (j > a) j = a;
      if (j > lw) j = lw;
      F[K] = j;
    }
    dl = (F[K] >= a) ? 0 : K + 1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int e = 1; e <= K; e++) {
      if (dl <= K) break;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int di = 0; di < ND; di++) {
        const int d = di - K;
        G[di] = LV_NONE;
        if (d < -e || d > e) continue;
        int best = LV_NONE;
        // substitution: (j, x) -> (j+1, x+1)
        if (d > -e && d < e && F[di] != LV_NONE) {
          const int j0 = F[di];
          best = (j0 < a && j0 + d < lw) ? j0 + 1 : j0;
        }
        // insertion (text only): from diagonal d-1, (j, x) -> (j, x+1)
        if (d - 1 >= -(e - 1) && F[di > 0 ? di - 1 : 0] != LV_NONE) {
          const int j0 = F[di > 0 ? di - 1 : 0];
          if (j0 + d - 1 < lw && j0 > best) best = j0;
        }
        // deletion (pattern only): from diagonal d+1, (j, x) -> (j+1, x); not at the wall, where
        // the last cell before it on that diagonal takes the step instead
        if (d + 1 <= e - 1 && F[di + 1 < ND ? di + 1 : ND - 1] != LV_NONE) {
          const int j0 = F[di + 1 < ND ? di + 1 : ND - 1];
          int cand = LV_NONE;
          if (j0 < a) {
            if (j0 + d + 1 < lw) cand = j0 + 1;
            else if (j0 - 1 >= (d + 1 < 0 ? -(d + 1) : 0)) cand = j0;   // from (j0-1, lw-1): lands on (j0, lw-1)
          }
          if (cand > best) best = cand;
        }
        if (best == LV_NONE) continue;
        if (best < a && best + d < lw) {
          int ext = lv_lce_bwd(mis[ND - 1 - di], a - 1 - best);   // forward diagonal -d
          if (best + ext > a) ext = a - best;
          if (best + ext + d > lw) ext = lw - d - best;
          if (ext > 0) best += ext;
        }
        G[di] = best;
        if (best >= a) dl = e;
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int i = 0; i < ND; i++) F[i] = G[i];
    }
    if (dl > K) return 0u;
  }

  // ---- right: pattern[a+L, m) against the text from the anchor's end ----------------------
  const int kr = K - dl;
  uint32_t ends = 0;
  const int j_start = a + L;
  {
    int F[ND], G[ND];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < ND; i++) F[i] = LV_NONE;
    {
      int j = j_start + lv_lce_fwd(mis[K], j_start);
      if (j > m) j = m;
      if (j > wr) j = wr;
      F[K] = j;
      if (j >= m) ends |= 1u << K;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int e = 1; e <= K; e++) {
      if (e > kr) break;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int di = 0; di < ND; di++) {
        const int d = di - K;
        G[di] = LV_NONE;
        if (d < -e || d > e) continue;
        int best = LV_NONE;
        if (d > -e && d < e && F[di] != LV_NONE) {      // substitution
          const int j0 = F[di];
          best = (j0 < m && j0 + d < wr) ? j0 + 1 : j0;
        }
        if (d - 1 >= -(e - 1) && F[di > 0 ? di - 1 : 0] != LV_NONE) {  // insertion: text character consumed
          const int j0 = F[di > 0 ? di - 1 : 0];
          if (j0 + d - 1 < wr && j0 > best) best = j0;
        }
        if (d + 1 <= e - 1 && F[di + 1 < ND ? di + 1 : ND - 1] != LV_NONE) {    // deletion: pattern character consumed
          const int j0 = F[di + 1 < ND ? di + 1 : ND - 1];
          // a diagonal that has reached the end of the pattern takes the step from its cell in
          // row m-1 (the last pattern character deleted), if it has one
          const int entry = j_start + (d + 1 < 0 ? -(d + 1) : 0);
          const int cand = (j0 < m) ? j0 + 1 : ((m - 1 >= entry) ? m : LV_NONE);
          if (cand > best) best = cand;
        }
        if (best == LV_NONE) continue;
        if (best < m && best + d < wr) {
          int ext = lv_lce_fwd(mis[di], best);
          if (best + ext > m) ext = m - best;
          if (best + ext + d > wr) ext = wr - d - best;
          if (ext > 0) best += ext;
        }
        G[di] = best;
        if (best >= m) ends |= 1u << di;
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int i = 0; i < ND; i++) F[i] = G[i];
    }
  }
  return ends;
}

}  // namespace pm
