// pm_index.cuh -- text-driven pair-seed index scan (sm_100a).
//
// The bit-plane kernels of pm_scan.cuh touch every (pattern, position) pair.
// For DNA-like alphabets that is far more work than the problem needs.  Cut
// the first P*L characters of every pattern into P = k+2 pieces of L
// characters.  k errors damage at most k pieces, so every occurrence with
// <= k errors holds two INTACT pieces i < j, and their text distance is
//   substitutions : exactly (j-i)*L
//   edits         : (j-i)*L - k .. (j-i)*L + k   (never less than L)
// A text position p therefore only has to ask, for each distance d of that
// small set D, "is (text[p..p+L), text[p+d..p+d+L)) a pair of some pattern?".
// This is the role the reference gives its keyword tree / hash table front
// ends (exact_bases.cc, exact_halves.cc, hash_table.cc), sharpened from one
// piece to two so that the filter passes ~1 % of the positions instead of ~70 %.
//
//   level 1  one bit per pair key in a 128 KB bitmap in SHARED memory:
//            idx = key * mult + salt[d];  word = idx >> (32-log_words),
//            bit = 31 - (idx & 31).  One LDS per (position, distance).
//   level 2  survivors go to a per-warp queue and are resolved 32 at a time
//            against the exact table in L2: bucket[hash2(key,d)] -> entries
//            (key, slot << 10 | d_index << 5 | piece offset); matches go to a
//            second queue and are verified 32 at a time.
//   verify   substitutions: XOR-free popcount on the packed window against the
//            pattern's per-code accept masks (IUPAC pattern symbols are a mask
//            with several codes set); edits: the reference's own recurrence
//            (index_verify_kernel).
//
// The text is kept 2 bits per character (+ 2 flag bits: "matches no pattern
// symbol" and "end of sequence"), 16 characters per 8-byte word.
//
// Exactness.  The candidate set is the one of pm_scan.cuh / the reference:
//   substitutions: the verifier counts the mismatches of the whole window and
//     rejects windows holding an end-of-sequence character -- the level-k hit
//     condition of shift_and_inexact.cc:253-326 without indels;
//   edits: the table only proposes seeds; the reference's recurrence decides.
// The pair lemma above makes the proposals a superset in both cases.
//
// Applicable when the text bytes fall into <= 4 match classes (+ "matches
// nothing"), no pattern is longer than 32 and L >= 4 (host checks); other
// inputs use the bit-plane kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pm_scan.cuh"
#include "pm_lv.cuh"

namespace pm {

constexpr int PAIR_THREADS = 1024;            // one CTA per SM: the bitmap fills its shared memory
constexpr int PAIR_WARPS = PAIR_THREADS / 32;
constexpr int PAIR_MAX_ND = 32;
constexpr int PAIR_LOG_WORDS = 15;            // 2^15 words = 128 KB = 2^20 bits
constexpr int PAIR_LMAX = 8;

struct IndexArgs {
  const uint2 *pk;          // [n16] x = 2-bit codes of 16 characters, y = flags (bit0 bad, bit1 eos)
  uint64_t n;               // shard length (characters)
  uint64_t n16;             // words in pk
  // level 1
  const uint32_t *bitmap;   // [1 << log_words]
  uint32_t log_words;
  uint32_t L;               // piece length
  uint32_t nd;              // distances
  uint32_t mult;            // odd multiplier of the level-1 hash
  uint32_t plain_patterns;  // every pattern character accepts exactly one text code (no IUPAC symbols)
  uint32_t l1_bij;          // level 1 indexed by (key * mult) mod 2^20, a bijection of keys of <= 20 bits
                            // (the edit kernels with L <= 5); else by the hash key * mult + salt
  uint32_t dist[PAIR_MAX_ND];
  uint32_t salt[PAIR_MAX_ND];
  // level 1b: a second bitmap keyed by (pair key, distance), consulted only for the positions
  // that passed level 1; log_words2 == 0: absent.  Edits: 64 KB in the shared memory level 1
  // leaves free.  Large pattern sets (level 1 saturates: beyond ~100 k entries): 16 MB in
  // global memory, L2-resident (b2_global).
  const uint32_t *bitmap2;
  uint32_t log_words2;
  uint32_t b2_global;
  uint32_t mult2;
  uint32_t salt2[PAIR_MAX_ND];
  // level 2
  const uint32_t *bucket;   // [(1 << log_buckets) + 1]
  uint32_t log_buckets;
  const uint2 *entries;     // x = key, y = slot << 10 | d_index << 5 | piece offset
  // patterns
  const ulonglong4 *pacc;   // [n_slots] bit 2j of .x/.y/.z/.w: character j accepts text code 0/1/2/3
  const uint8_t *plen;      // [n_slots]
  const uint32_t *umask4;   // [n_slots][4] Shift-And masks per text code (edit verifier)
  unsigned long long *out;  // candidates: slot << 40 | pos
  unsigned long long *count;
  uint64_t cap;
  int64_t lo, hi;           // emit candidates with lo <= pos < hi
  int64_t p_begin, p_end;   // first-piece positions handled by this launch
  uint32_t k;
  // edit mode: seeds instead of candidates
  unsigned long long *seeds, *seed_n, *seed_overflow, *seed_count;
  uint64_t seed_cap;
  unsigned long long *work_count;  // level-1 survivor records (word, distance) sent to stage A
};

// 16-character words of the packed text.  Characters past the end read as end
// of sequence (the pack kernel fills whole tiles).
__device__ __forceinline__ uint2 pk_load(const IndexArgs &a, int64_t w) {
  if (w < 0 || (uint64_t)w >= a.n16) return make_uint2(0u, 0xffffffffu);
  return __ldg(a.pk + w);
}
__device__ __forceinline__ uint32_t pk_codes(const IndexArgs &a, int64_t w) {
  if (w < 0 || (uint64_t)w >= a.n16) return 0u;
  return __ldg(&a.pk[w].x);
}

// codes and flags of the 32 characters starting at s
__device__ __forceinline__ void pk_window(const IndexArgs &a, int64_t s, uint64_t &codes, uint64_t &flags) {
  const int64_t w = s >> 4;  // arithmetic shift: s may be negative near the start
  const uint32_t sh = 2u * (uint32_t)(s & 15);
  const uint2 a0 = pk_load(a, w), a1 = pk_load(a, w + 1), a2 = pk_load(a, w + 2);
  const uint32_t c0 = __funnelshift_r(a0.x, a1.x, sh), c1 = __funnelshift_r(a1.x, a2.x, sh);
  const uint32_t f0 = __funnelshift_r(a0.y, a1.y, sh), f1 = __funnelshift_r(a1.y, a2.y, sh);
  codes = (uint64_t)c0 | ((uint64_t)c1 << 32);
  flags = (uint64_t)f0 | ((uint64_t)f1 << 32);
}

__device__ __forceinline__ void index_emit(const IndexArgs &a, uint32_t slot, int64_t pos) {
  if (pos < a.lo || pos >= a.hi) return;
  const unsigned long long at = atomicAdd(a.count, 1ULL);
  if (at < a.cap) a.out[at] = ((unsigned long long)slot << 40) | (unsigned long long)pos;
}

__host__ __device__ __forceinline__ uint32_t pair_hash1(uint32_t key, uint32_t mult, uint32_t salt) {
  return key * mult + salt;
}
__host__ __device__ __forceinline__ uint32_t pair_hash2(uint32_t key, uint32_t di, uint32_t log_buckets) {
  return ((key + di * 0x9E3779B1u) * 0x85EBCA6Bu) >> (32u - log_buckets);
}

// one slot of the candidate list for every lane that has one: a single atomic per warp (all 32 lanes
// call this together)
__device__ __forceinline__ void index_emit_warp(const IndexArgs &a, bool has, uint32_t slot, int64_t pos) {
  has = has && pos >= a.lo && pos < a.hi;
  const uint32_t bal = __ballot_sync(0xffffffffu, has);
  if (bal == 0u) return;
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long base = 0;
  if (lane == (uint32_t)(__ffs(bal) - 1)) base = atomicAdd(a.count, (unsigned long long)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
  if (has) {
    const unsigned long long at = base + __popc(bal & ((1u << lane) - 1u));
    if (at < a.cap) a.out[at] = ((unsigned long long)slot << 40) | (unsigned long long)pos;
  }
}

// substitutions: mismatches of the whole window of `slot` starting at s; true: a candidate at *pos
__device__ __forceinline__ bool pair_verify_window(const IndexArgs &a, uint32_t slot, int64_t s, int64_t *pos) {
  const uint32_t m = a.plen[slot] & 0x7fu;
  if (s < 0 || (uint64_t)(s + m) > a.n) return false;
  uint64_t codes, flags;
  pk_window(a, s, codes, flags);
  const ulonglong2 *accp = reinterpret_cast<const ulonglong2 *>(a.pacc + slot);
  const ulonglong2 acc01 = __ldg(accp), acc23 = __ldg(accp + 1);
  ulonglong4 acc;
  acc.x = acc01.x; acc.y = acc01.y; acc.z = acc23.x; acc.w = acc23.y;
  constexpr uint64_t M5 = 0x5555555555555555ULL;
  const uint64_t lenmask = (m >= 32) ? ~0ULL : ((1ULL << (2 * m)) - 1ULL);
  const uint64_t lo = codes & M5, hi = (codes >> 1) & M5;
  const uint64_t match = (~(hi | lo) & acc.x) | (lo & ~hi & acc.y) | (hi & ~lo & acc.z) | (hi & lo & acc.w);
  const uint64_t mis = (~match | flags) & M5 & lenmask;
  *pos = s + m;
  return __popcll(mis) <= (int)a.k && (flags & ~M5 & lenmask) == 0;
}

// Level 2 and verification run as two warp-wide stages fed from per-warp queues in
// shared memory, so that the rare work is done with (nearly) all 32 lanes busy:
//   stage A  one record per lane = (packed word, distance index, 16-bit mask of the
//            positions of that word whose level-1 bit was set): recompute the pair
//            key of each set position, walk its bucket of the exact table;
//   stage B  substitutions only: one (window start, slot) per lane, mismatch count.
// Both queues belong to one warp and are only touched by warp-collective code, so
// their fill counts live in (warp-uniform) registers and there are no atomics.
constexpr int PAIR_QCAP = 64;  // entries per warp and queue: < 32 before a push of <= 32

struct PairQueues {
  unsigned long long *qa, *qb;  // this warp's queues
  uint32_t na, nb;              // fill counts, warp-uniform
};

template <bool EDIT>
__device__ __forceinline__ void pair_stage_b(const IndexArgs &a, bool valid, unsigned long long rec) {
  if (EDIT) return;
  // (called by all 32 lanes: candidates are appended with one atomic per warp)
  const uint32_t slot = (uint32_t)(rec & 0x3fffffu);
  int64_t pos = 0;
  const bool hit = valid && pair_verify_window(a, slot, (int64_t)(rec >> 22) - SEED_BIAS, &pos);
  index_emit_warp(a, hit, slot, pos);
}

template <bool EDIT>
__device__ __forceinline__ void pair_stage_a(const IndexArgs &a, PairQueues &q, bool valid,
                                             unsigned long long rec, uint32_t lane) {
  const int64_t w = (int64_t)(rec >> 21);
  const uint32_t di = (uint32_t)(rec >> 16) & 31u;
  uint32_t hm = valid ? (uint32_t)(rec & 0xffffu) : 0u;
  const uint32_t L2 = 2u * a.L;
  const uint32_t maskL = (1u << L2) - 1u;
  const uint32_t d2 = 2u * a.dist[valid ? di : 0u];
  uint32_t c0 = 0, c1 = 0, c2 = 0;
  if (valid) { c0 = pk_codes(a, w); c1 = pk_codes(a, w + 1); c2 = pk_codes(a, w + 2); }
  uint32_t e0 = 0, e1 = 0, key = 0;
  int64_t p = 0;
  const uint32_t lt = (1u << lane) - 1u;
  for (;;) {
    if (!__any_sync(0xffffffffu, (hm != 0u) || (e0 < e1))) break;
    if (e0 >= e1 && hm) {  // next set position of the word
      const int b = __ffs(hm) - 1;
      hm &= hm - 1;
      const uint32_t i2 = 2u * (uint32_t)(15 - b);
      const uint32_t v0 = __funnelshift_r(c0, c1, i2), v1 = __funnelshift_r(c1, c2, i2), v2 = c2 >> i2;
      const uint32_t hi = (d2 >= 32u) ? __funnelshift_r(v1, v2, d2) : __funnelshift_r(v0, v1, d2);
      key = (v0 & maskL) | ((hi & maskL) << L2);
      const uint32_t h = pair_hash2(key, di, a.log_buckets);
      e0 = __ldg(a.bucket + h);
      e1 = __ldg(a.bucket + h + 1);
      p = (w << 4) + (15 - b);
    }
    bool found = false;
    uint32_t meta = 0;
    if (e0 < e1) {
      const uint2 ent = __ldg(a.entries + e0);
      ++e0;
      meta = ent.y;
      found = (ent.x == key) && (((meta >> 5) & 31u) == di);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, found);
    if (bal == 0u) continue;
    const uint32_t slot = meta >> 10;
    const int64_t s = p - (int64_t)(meta & 31u) + SEED_BIAS;  // nominal window start, biased
    const uint32_t rank = __popc(bal & lt), cnt = __popc(bal);
    if (EDIT) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.seed_n, (unsigned long long)cnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (found) {
        const unsigned long long at = base + rank;
        if (at < a.seed_cap) a.seeds[at] = ((unsigned long long)slot << 40) | (unsigned long long)s;
        else *a.seed_overflow = 1ULL;
      }
    } else {
      if (found) q.qb[q.nb + rank] = ((unsigned long long)s << 22) | slot;
      q.nb += cnt;
      if (q.nb >= 32u) {
        __syncwarp();
        const unsigned long long r = q.qb[q.nb - 32u + lane];
        __syncwarp();
        q.nb -= 32u;
        pair_stage_b<EDIT>(a, true, r);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Pieces shared by the three scan kernels below.  One persistent 1024-thread CTA per
// SM; a warp owns chunks of 32 packed words (512 positions): lane l holds word l and
// gets the next two words by shuffle, so the 16 positions of its word see their
// 32-character context without any further load.
struct PairCta {
  uint32_t *bitmap;          // shared-memory copy of level 1
  const uint32_t *bitmap2;   // level 1b: shared-memory copy, or the global table (b2_global)
  uint32_t n_words2;
  PairQueues q;
  uint32_t lane, lt;
  int64_t w_begin, n_chunks;
  unsigned long long n_queued;
};

__device__ __forceinline__ void pair_cta_init(const IndexArgs &a, uint32_t *s_mem, PairCta &t) {
  const uint32_t n_words = 1u << a.log_words;
  t.bitmap = s_mem;
  unsigned long long *s_queue = reinterpret_cast<unsigned long long *>(s_mem + n_words);
  t.n_words2 = a.bitmap2 ? (1u << a.log_words2) : 0u;
  uint32_t *s_bitmap2 = s_mem + n_words + PAIR_WARPS * 2 * PAIR_QCAP * 2;  // behind the queues
  t.bitmap2 = a.b2_global ? a.bitmap2 : s_bitmap2;
  for (uint32_t i = threadIdx.x; i < n_words / 4; i += blockDim.x)
    reinterpret_cast<uint4 *>(t.bitmap)[i] = __ldg(reinterpret_cast<const uint4 *>(a.bitmap) + i);
  if (!a.b2_global)
    for (uint32_t i = threadIdx.x; i < t.n_words2 / 4; i += blockDim.x)
      reinterpret_cast<uint4 *>(s_bitmap2)[i] = __ldg(reinterpret_cast<const uint4 *>(a.bitmap2) + i);
  __syncthreads();
  t.lane = threadIdx.x & 31;
  t.lt = (1u << t.lane) - 1u;
  const uint32_t warp = threadIdx.x >> 5;
  t.q.qa = s_queue + (size_t)warp * (2 * PAIR_QCAP);
  t.q.qb = t.q.qa + PAIR_QCAP;
  t.q.na = t.q.nb = 0;
  t.w_begin = a.p_begin >> 4;
  t.n_chunks = (((a.p_end + 15) >> 4) - t.w_begin + 31) >> 5;
  t.n_queued = 0;
}

// codes of this lane's word and the two after it; mask of the word's positions that lie
// inside [p_begin, p_end) (position i <-> bit 15-i)
__device__ __forceinline__ void pair_load_words(const IndexArgs &a, const PairCta &t, int64_t chunk,
                                                int64_t &w, uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                                uint32_t &vm) {
  const int64_t wc0 = t.w_begin + (chunk << 5);
  w = wc0 + t.lane;
  const uint32_t mine = pk_codes(a, w);
  const uint32_t extra = (t.lane < 2) ? pk_codes(a, wc0 + 32 + t.lane) : 0u;
  const uint32_t x0 = __shfl_sync(0xffffffffu, extra, 0), x1 = __shfl_sync(0xffffffffu, extra, 1);
  c0 = mine;
  c1 = __shfl_down_sync(0xffffffffu, mine, 1);
  c2 = __shfl_down_sync(0xffffffffu, mine, 2);
  if (t.lane == 31) { c1 = x0; c2 = x1; }
  if (t.lane == 30) c2 = x0;
  const int64_t base = w << 4;
  vm = 0xffffu;
  if (base < a.p_begin) vm &= (a.p_begin - base >= 16) ? 0u : (0xffffu >> (uint32_t)(a.p_begin - base));
  if (base + 16 > a.p_end) vm &= (a.p_end <= base) ? 0u : (0xffffu << (uint32_t)(base + 16 - a.p_end));
}

// the words of a later chunk towards L2 (no register is held: the kernels sit at their register cap)
__device__ __forceinline__ void pair_prefetch_words(const IndexArgs &a, const PairCta &t, int64_t chunk) {
  const int64_t w = t.w_begin + (chunk << 5) + t.lane;
  if (chunk < t.n_chunks && (uint64_t)w < a.n16) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pk + w));
}

// level 1b (edits): the shared bits of level 1 say "some distance"; keep the positions whose
// (key, distance) bit is set as well.  Divergent, only for lanes with a hit.
__device__ __forceinline__ uint32_t pair_level1b(const IndexArgs &a, const PairCta &t, uint32_t hm, uint32_t di,
                                                 uint32_t c0, uint32_t c1, uint32_t c2) {
  const uint32_t L2 = 2u * a.L, maskL = (1u << L2) - 1u;
  const uint32_t d2 = 2u * a.dist[di], salt2 = a.salt2[di], wsh2 = 32u - a.log_words2;
  uint32_t keep = 0u;
  while (hm) {
    const int b = __ffs(hm) - 1;
    hm &= hm - 1;
    const uint32_t i2 = 2u * (uint32_t)(15 - b);
    const uint32_t v0 = __funnelshift_r(c0, c1, i2), v1 = __funnelshift_r(c1, c2, i2), v2 = c2 >> i2;
    const uint32_t hi = (d2 >= 32u) ? __funnelshift_r(v1, v2, d2) : __funnelshift_r(v0, v1, d2);
    const uint32_t key = (v0 & maskL) | ((hi & maskL) << L2);
    const uint32_t idx2 = key * a.mult2 + salt2;
    const uint32_t wrd = a.b2_global ? __ldg(a.bitmap2 + (idx2 >> wsh2)) : t.bitmap2[idx2 >> wsh2];
    if ((wrd >> (idx2 & 31u)) & 1u) keep |= 1u << b;
  }
  return keep;
}

// queue the (word, distance, position mask) records of the warp; run stage A when 32 wait
template <bool EDIT>
__device__ __forceinline__ void pair_push(const IndexArgs &a, PairCta &t, int64_t w, uint32_t di, uint32_t hm) {
  const uint32_t bal = __ballot_sync(0xffffffffu, hm != 0u);
  if (bal == 0u) return;
  if (hm) t.q.qa[t.q.na + __popc(bal & t.lt)] = ((unsigned long long)w << 21) | (di << 16) | hm;
  t.q.na += __popc(bal);
  t.n_queued += __popc(bal);
  if (t.q.na >= 32u) {
    __syncwarp();
    const unsigned long long rec = t.q.qa[t.q.na - 32u + t.lane];
    __syncwarp();
    t.q.na -= 32u;
    pair_stage_a<EDIT>(a, t.q, true, rec, t.lane);
  }
}

template <bool EDIT>
__device__ __forceinline__ void pair_cta_finish(const IndexArgs &a, PairCta &t) {
  __syncwarp();
  {
    const bool valid = t.lane < t.q.na;
    const unsigned long long rec = valid ? t.q.qa[t.lane] : 0ULL;
    __syncwarp();
    t.q.na = 0;
    pair_stage_a<EDIT>(a, t.q, valid, rec, t.lane);
  }
  __syncwarp();
  if (!EDIT) {
    const bool valid = t.lane < t.q.nb;
    const unsigned long long rec = valid ? t.q.qb[t.lane] : 0ULL;
    pair_stage_b<EDIT>(a, valid, rec);
  }
  if (t.lane == 0 && a.work_count) atomicAdd(a.work_count, t.n_queued);
}

// one level-1 lookup: bit 31-(idx&31) of the bitmap word -> shifted into bit 0 of hm
__device__ __forceinline__ uint32_t pair_probe(const uint32_t *bitmap, uint32_t key, uint32_t mult, uint32_t salt,
                                               uint32_t wsh, uint32_t hm) {
  const uint32_t idx = key * mult + salt;
  const uint32_t wrd = bitmap[idx >> wsh];
  return __funnelshift_l(__funnelshift_l(wrd, wrd, idx), hm, 1);
}

// the same through the bijective index of short keys (IndexArgs::l1_bij): word = top 15 bits of
// (key * mult mod 2^20), bit 31 - (key & 31).  For a fixed value of the key's low five bits the
// word index is a bijection of the other fifteen (odd multiplier), so no two keys share a bit;
// the rotation takes the key itself: LOP3 key, IMAD, SHF + LOP3 address, LDS, 2 x SHF per lookup.
__device__ __forceinline__ uint32_t pair_probe_bij(uint32_t bitmap_saddr, uint32_t key, uint32_t mult_hi, uint32_t hm) {
  // address = base + 4 * word as ONE multiply-add (FMA pipe; the ALU pipe is the one that binds):
  // written in PTX because the compiler turns the C expression into shift + mask + add
  uint32_t addr, wrd;
  asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"((key * mult_hi) >> 17), "r"(bitmap_saddr));
  asm("ld.shared.u32 %0, [%1];" : "=r"(wrd) : "r"(addr));
  return __funnelshift_l(__funnelshift_l(wrd, wrd, key), hm, 1);
}

// Generic kernel: run-time piece length and distance set, one distance after the other.
// Serves the edit scans (13-19 distances: they are bound by the shared-memory wavefront
// rate of the lookups themselves -- a variant sharing the shifted windows between the
// distances of a group, as the substitution kernel below does, measured 5 % slower).
template <bool EDIT>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
pair_scan_kernel(const __grid_constant__ IndexArgs a) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  PairCta t;
  pair_cta_init(a, s_mem, t);
  const uint32_t wsh = 32u - a.log_words;
  const uint32_t L2 = 2u * a.L;
  const uint32_t maskL = (1u << L2) - 1u;       // L <= 8
  const uint32_t maskHi = maskL << L2;
  const uint32_t mult = a.mult;
  const uint32_t warp = threadIdx.x >> 5;

  for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; c < t.n_chunks; c += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    int64_t w;
    uint32_t c0, c1, c2, vm;
    pair_load_words(a, t, c, w, c0, c1, c2, vm);
    uint32_t av[16];
#pragma unroll
    for (int i = 0; i < 16; i++) av[i] = __funnelshift_r(c0, c1, 2 * i) & maskL;

    for (uint32_t di = 0; di < a.nd; ++di) {
      const uint32_t s = 2u * (a.dist[di] - a.L);  // < 64
      const uint32_t salt = a.salt[di];
      const bool hiw = s >= 32u;
      const uint32_t y0 = hiw ? c1 : c0, y1 = hiw ? c2 : c1, y2 = hiw ? 0u : c2;
      const uint32_t X0 = __funnelshift_r(y0, y1, s), X1 = __funnelshift_r(y1, y2, s);  // shift taken mod 32
      uint32_t hm = 0;
#pragma unroll
      for (int i = 0; i < 16; i++)
        hm = pair_probe(t.bitmap, (__funnelshift_r(X0, X1, 2 * i) & maskHi) | av[i], mult, salt, wsh, hm);
      hm &= vm;
      if (t.n_words2 && hm) hm = pair_level1b(a, t, hm, di, c0, c1, c2);
      pair_push<EDIT>(a, t, w, di, hm);
    }
  }
  pair_cta_finish<EDIT>(a, t);
}

typedef void (*PairKernel)(const IndexArgs);

// ---------------------------------------------------------------------------
// Edit scan with compile-time piece length and error count.  The distances a piece pair can
// have are known at compile time (EditDist), so the lookup loop is straight-line code over
// shifted windows shared between positions and distances (block by block, to stay inside 64
// registers): 7 instructions per lookup instead of 10.  Level 1b is not asked inline -- there
// it ran as a divergent per-lane loop with 9 of 32 lanes busy, a third of the old kernel's
// instructions (profiles/README.md, r2_before_edit25_scan) -- but by stage A, one queued
// record per lane.  Seeds carry the index of their first piece: index_verify_lv_kernel
// extends from that anchor (pm_lv.cuh).
template <int L, int K>
struct EditDist {
  static constexpr int P = K + 2;
  static constexpr int DMAX = (P - 1) * L + K;
  __host__ __device__ static constexpr bool used(int d) {
    if (d < L || d > DMAX) return false;
    for (int g = 1; g < P; g++)
      if (d >= g * L - K && d <= g * L + K) return true;
    return false;
  }
  // position of d in the ascending list of used distances (IndexArgs::dist)
  __host__ __device__ static constexpr int index(int d) {
    int n = 0;
    for (int x = L; x < d; x++) n += used(x) ? 1 : 0;
    return n;
  }
};

constexpr int SEED_POS_BITS = 39;   // edit seeds of the index path: slot << 42 | piece index << 39 | biased start

// Stage A of the edit scan, in two steps so that neither waits for the other's stragglers:
//   A1  one queued record per lane = (lane that owns the word, distance index, 16-bit mask of the
//       word's positions that passed level 1).  Every set position asks level 1b (the shared-memory
//       bitmap keyed by (pair key, distance)); what passes is compacted across the warp into a
//       32-entry staging row (lane, position, distance);
//   A2  32 staged entries at a time: one (position, distance) per lane walks its bucket of the
//       exact table in L2; matches become seeds (slot, piece index, nominal start).
// The text words of a record's lane come from that lane's registers (shuffle), not from memory.
template <int L>
__device__ __forceinline__ uint32_t pair_edit_key(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t b, uint32_t d2) {
  constexpr uint32_t L2 = 2u * L;
  constexpr uint32_t maskL = (1u << L2) - 1u;
  const uint32_t i2 = 2u * (15u - b);
  const uint32_t v0 = __funnelshift_r(c0, c1, i2), v1 = __funnelshift_r(c1, c2, i2), v2 = c2 >> i2;
  const uint32_t hi = (d2 >= 32u) ? __funnelshift_r(v1, v2, d2) : __funnelshift_r(v0, v1, d2);
  return (v0 & maskL) | ((hi & maskL) << L2);
}

template <int L>
__device__ __forceinline__ void pair_edit_a2(const IndexArgs &a, int64_t w_base, uint32_t cw0, uint32_t cw1,
                                             uint32_t cw2, bool valid, uint32_t ent, uint32_t lane) {
  const uint32_t src = (ent >> 9) & 31u, b = (ent >> 5) & 15u, di = ent & 31u;
  const uint32_t c0 = __shfl_sync(0xffffffffu, cw0, src), c1 = __shfl_sync(0xffffffffu, cw1, src),
                 c2 = __shfl_sync(0xffffffffu, cw2, src);
  uint32_t e0 = 0, e1 = 0, key = 0;
  if (valid) {
    key = pair_edit_key<L>(c0, c1, c2, b, 2u * a.dist[di]);
    const uint32_t h = pair_hash2(key, di, a.log_buckets);
    e0 = __ldg(a.bucket + h);
    e1 = __ldg(a.bucket + h + 1);
  }
  const int64_t p = ((w_base + (int64_t)src) << 4) + (15 - (int)b);
  const uint32_t lt = (1u << lane) - 1u;
  while (__any_sync(0xffffffffu, e0 < e1)) {
    bool found = false;
    uint32_t meta = 0;
    if (e0 < e1) {
      const uint2 en = __ldg(a.entries + e0);
      ++e0;
      meta = en.y;
      found = (en.x == key) && (((meta >> 5) & 31u) == di);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, found);
    if (bal == 0u) continue;
    const uint32_t slot = meta >> 10, off = meta & 31u;
    const int64_t s = p - (int64_t)off + SEED_BIAS;  // nominal window start, biased
    const uint32_t rank = __popc(bal & lt), cnt = __popc(bal);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.seed_n, (unsigned long long)cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (found) {
      const unsigned long long at = base + rank;
      if (at < a.seed_cap)
        a.seeds[at] = ((unsigned long long)slot << (SEED_POS_BITS + 3)) | ((unsigned long long)(off / L) << SEED_POS_BITS) |
                      (unsigned long long)s;
      else *a.seed_overflow = 1ULL;
    }
  }
}

// A1 over 32 records (one per lane); `stage` = this warp's 32 staging entries, n_staged how many
// wait there (warp-uniform, < 32 on entry and on return).
template <int L>
__device__ __forceinline__ void pair_edit_a1(const IndexArgs &a, const uint32_t *bitmap2, uint16_t *stage,
                                             uint32_t &n_staged, int64_t w_base, uint32_t cw0, uint32_t cw1, uint32_t cw2,
                                             bool valid, uint32_t rec, uint32_t lane) {
  const uint32_t src = (rec >> 21) & 31u, di = (rec >> 16) & 31u;
  uint32_t hm = valid ? (rec & 0xffffu) : 0u;
  const uint32_t c0 = __shfl_sync(0xffffffffu, cw0, src), c1 = __shfl_sync(0xffffffffu, cw1, src),
                 c2 = __shfl_sync(0xffffffffu, cw2, src);
  const uint32_t d2 = 2u * a.dist[valid ? di : 0u];
  const uint32_t salt2 = a.salt2[valid ? di : 0u], wsh2 = 32u - a.log_words2;
  const uint32_t lt = (1u << lane) - 1u;
  while (__any_sync(0xffffffffu, hm != 0u)) {
    bool pass = false;
    uint32_t ent = 0;
    if (hm) {
      const uint32_t b = (uint32_t)(__ffs(hm) - 1);
      hm &= hm - 1;
      const uint32_t idx2 = pair_edit_key<L>(c0, c1, c2, b, d2) * a.mult2 + salt2;
      pass = (bitmap2[idx2 >> wsh2] >> (idx2 & 31u)) & 1u;
      ent = (src << 9) | (b << 5) | di;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, pass);
    if (bal == 0u) continue;
    const uint32_t at = n_staged + __popc(bal & lt), total = n_staged + __popc(bal);
    if (pass && at < 32u) stage[at] = (uint16_t)ent;
    if (total >= 32u) {     // a full row: A2 takes it, the rest of this round starts the next row
      __syncwarp();
      const uint32_t e = stage[lane];
      __syncwarp();
      pair_edit_a2<L>(a, w_base, cw0, cw1, cw2, true, e, lane);
      if (pass && at >= 32u) stage[at - 32u] = (uint16_t)ent;
      n_staged = total - 32u;
    } else {
      n_staged = total;
    }
  }
}

constexpr int EDIT_DBLOCK = 8;     // distances handled per block of shifted windows

constexpr int EDIT_QCAP = 256;     // 4-byte records per warp: the queues' kilobyte per warp

// One chunk the slow way (run-time distance loop, stage A after every distance): only for a chunk
// whose records did not fit the queue -- text that repeats a piece pair of some pattern over and over.
template <int L>
__device__ __noinline__ uint32_t pair_edit_chunk_slow(const IndexArgs &a, const uint32_t *bitmap, const uint32_t *bitmap2,
                                                      uint16_t *stage, int64_t w_base, uint32_t c0, uint32_t c1,
                                                      uint32_t c2, uint32_t vm, uint32_t lane) {
  uint32_t n_queued = 0, n_staged = 0;
  const uint32_t wsh = 32u - a.log_words;
  constexpr uint32_t maskL = (1u << (2 * L)) - 1u, maskHi = maskL << (2 * L);
  for (uint32_t di = 0; di < a.nd; ++di) {
    const uint32_t s = 2u * (a.dist[di] - L);  // < 64
    const bool hiw = s >= 32u;
    const uint32_t y0 = hiw ? c1 : c0, y1 = hiw ? c2 : c1, y2 = hiw ? 0u : c2;
    const uint32_t X0 = __funnelshift_r(y0, y1, s), X1 = __funnelshift_r(y1, y2, s);
    uint32_t hm = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t key = (__funnelshift_r(X0, X1, 2 * i) & maskHi) | (__funnelshift_r(c0, c1, 2 * i) & maskL);
      hm = (4 * L <= 20) ? pair_probe_bij((uint32_t)__cvta_generic_to_shared(bitmap), key, a.mult << 12, hm) : pair_probe(bitmap, key, a.mult, 0u, wsh, hm);
    }
    hm &= vm;
    if (__any_sync(0xffffffffu, hm != 0u)) {
      n_queued += __popc(__ballot_sync(0xffffffffu, hm != 0u));
      pair_edit_a1<L>(a, bitmap2, stage, n_staged, w_base, c0, c1, c2, hm != 0u, (lane << 21) | (di << 16) | hm, lane);
    }
  }
  __syncwarp();
  const uint32_t e = (lane < n_staged) ? stage[lane] : 0u;
  __syncwarp();
  pair_edit_a2<L>(a, w_base, c0, c1, c2, lane < n_staged, e, lane);
  return n_queued;
}

template <int L, int K>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
pair_scan_edit_kernel(const __grid_constant__ IndexArgs a) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  using ED = EditDist<L, K>;
  PairCta t;
  pair_cta_init(a, s_mem, t);
  const uint32_t wsh = 32u - a.log_words;
  constexpr uint32_t maskL = (1u << (2 * L)) - 1u;
  constexpr uint32_t mask2L = (L >= 8) ? 0xffffffffu : ((1u << (4 * L)) - 1u);
  constexpr uint32_t maskHi = mask2L & ~maskL;
  constexpr int NB = (ED::DMAX - L + 1 + EDIT_DBLOCK - 1) / EDIT_DBLOCK;
  const uint32_t mult = a.mult;
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t *queue = reinterpret_cast<uint32_t *>(t.q.qa);   // this warp's kilobyte: EDIT_QCAP records
  // the staging rows of A1 / A2 lie behind level 1b (the host sizes the shared memory for them)
  uint16_t *stage = reinterpret_cast<uint16_t *>(s_mem + (1u << a.log_words) + PAIR_WARPS * 2 * PAIR_QCAP * 2 +
                                                 (a.b2_global ? 0u : t.n_words2)) + warp * 32;
  const uint32_t lane_tag = t.lane << 21;
  const uint32_t bitmap_saddr = (uint32_t)__cvta_generic_to_shared(t.bitmap);

  for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; c < t.n_chunks; c += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    int64_t w;
    uint32_t cw[5], vm;
    pair_load_words(a, t, c, w, cw[0], cw[1], cw[2], vm);
    pair_prefetch_words(a, t, c + (int64_t)gridDim.x * (blockDim.x >> 5));
    cw[3] = cw[4] = 0u;
    uint32_t av[16];
#pragma unroll
    for (int i = 0; i < 16; i++) av[i] = __funnelshift_r(cw[0], cw[1], 2 * i) & maskL;
    uint32_t na = 0;        // records queued for this chunk (warp-uniform); > EDIT_QCAP: overflow
#pragma unroll
    for (int b = 0; b < NB; b++) {
      // windows G[tt] = text from character (d_lo - L + tt): bits [2L, 4L) hold the piece at distance
      // d_lo + (tt - i) from position i
      const int d_lo = L + b * EDIT_DBLOCK;
      uint32_t G[16 + EDIT_DBLOCK - 1];
#pragma unroll
      for (int tt = 0; tt < 16 + EDIT_DBLOCK - 1; tt++) {
        const int q = d_lo - L + tt;
        G[tt] = (q < 48) ? __funnelshift_r(cw[q >> 4], cw[(q >> 4) + 1], 2 * (q & 15)) : 0u;
      }
#pragma unroll
      for (int dd = 0; dd < EDIT_DBLOCK; dd++) {
        const int d = d_lo + dd;
        if (!ED::used(d)) continue;
        uint32_t hm = 0u;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const uint32_t key = (G[i + dd] & maskHi) | av[i];
          hm = (4 * L <= 20) ? pair_probe_bij(bitmap_saddr, key, mult << 12, hm) : pair_probe(t.bitmap, key, mult, 0u, wsh, hm);
        }
        hm &= vm;
        // queue (lane, distance, positions): no stage A here, the lookups stay straight-line code
        const uint32_t bal = __ballot_sync(0xffffffffu, hm != 0u);
        const uint32_t at = na + __popc(bal & t.lt);
        if (hm && at < (uint32_t)EDIT_QCAP) queue[at] = lane_tag | ((uint32_t)ED::index(d) << 16) | hm;
        na += __popc(bal);
      }
    }
    const int64_t w_base = w - t.lane;
    if (na > (uint32_t)EDIT_QCAP) {
      t.n_queued += pair_edit_chunk_slow<L>(a, t.bitmap, t.bitmap2, stage, w_base, cw[0], cw[1], cw[2], vm, t.lane);
    } else {
      t.n_queued += na;
      __syncwarp();
      uint32_t n_staged = 0;
      for (uint32_t r0 = 0; r0 < na; r0 += 32u) {
        const bool valid = r0 + t.lane < na;
        const uint32_t rec = valid ? queue[r0 + t.lane] : 0u;
        pair_edit_a1<L>(a, t.bitmap2, stage, n_staged, w_base, cw[0], cw[1], cw[2], valid, rec, t.lane);
      }
      __syncwarp();
      const uint32_t e = (t.lane < n_staged) ? stage[t.lane] : 0u;
      __syncwarp();
      pair_edit_a2<L>(a, w_base, cw[0], cw[1], cw[2], t.lane < n_staged, e, t.lane);
    }
  }
  if (t.lane == 0 && a.work_count) atomicAdd(a.work_count, t.n_queued);
}

template <int K>
inline PairKernel pair_edit_kernel_l(int L) {
  switch (L) {
    case 4: return pair_scan_edit_kernel<4, K>;
    case 5: return pair_scan_edit_kernel<5, K>;
    case 6: if (K <= 2) return pair_scan_edit_kernel<6, (K <= 2 ? K : 1)>; break;
    case 7: if (K <= 2) return pair_scan_edit_kernel<7, (K <= 2 ? K : 1)>; break;
    case 8: if (K <= 1) return pair_scan_edit_kernel<8, 1>; break;
  }
  return nullptr;
}
// specialised edit kernel for (L, K), or nullptr
inline PairKernel pair_edit_kernel(int L, int K) {
  switch (K) {
    case 1: return pair_edit_kernel_l<1>(L);
    case 2: return pair_edit_kernel_l<2>(L);
    case 3: return pair_edit_kernel_l<3>(L);
  }
  return nullptr;
}

// Edit verifier of pair_scan_edit_kernel's seeds (pm_lv.cuh).  Every seed goes through the q-gram
// filter, one seed per lane; the few that pass wait in a per-warp queue until 32 are there and
// are then extended from their first piece with all lanes busy (a warp running the extension for
// one lane in 32 was half of this kernel's instructions, profiles/README.md).
__device__ __forceinline__ void lv_load_seed(const IndexArgs &a, unsigned long long sd, int K, LvCtx &t, uint32_t &slot,
                                             int &pi, int64_t &s_nom) {
  slot = (uint32_t)(sd >> (SEED_POS_BITS + 3));
  pi = (int)((sd >> SEED_POS_BITS) & 7ULL);
  s_nom = (int64_t)(sd & ((1ULL << SEED_POS_BITS) - 1ULL)) - SEED_BIAS;
  const ulonglong2 *accp = reinterpret_cast<const ulonglong2 *>(a.pacc + slot);
  const ulonglong2 acc01 = __ldg(accp), acc23 = __ldg(accp + 1);
  lv_set_pattern(t, (int)(a.plen[slot] & 0x7fu), acc01.x, acc01.y, acc23.x, acc23.y, (int)a.plain_patterns);
  const int64_t s0 = s_nom - K;
  const int64_t wb = s0 >> 4;   // arithmetic shift: s0 may be negative near the start
  const uint32_t sh = 2u * (uint32_t)(s0 & 15);
  const uint2 p0 = pk_load(a, wb), p1 = pk_load(a, wb + 1), p2 = pk_load(a, wb + 2), p3 = pk_load(a, wb + 3);
  t.c[0] = __funnelshift_r(p0.x, p1.x, sh); t.c[1] = __funnelshift_r(p1.x, p2.x, sh); t.c[2] = __funnelshift_r(p2.x, p3.x, sh);
  t.f[0] = __funnelshift_r(p0.y, p1.y, sh); t.f[1] = __funnelshift_r(p1.y, p2.y, sh); t.f[2] = __funnelshift_r(p2.y, p3.y, sh);
}

template <int K>
__device__ __forceinline__ void lv_extend_seed(const IndexArgs &a, bool valid, unsigned long long sd) {
  if (!valid) return;
  LvCtx t;
  uint32_t slot;
  int pi;
  int64_t s_nom;
  lv_load_seed(a, sd, K, t, slot, pi, s_nom);
  const uint32_t ends = lv_verify<K>(t, pi * (int)a.L, (int)a.L);
#pragma unroll
  for (int d = -K; d <= K; d++)
    if ((ends >> (d + K)) & 1u) index_emit(a, slot, s_nom + t.m + d);
}

constexpr int VERIFY_LV_THREADS = 128;
template <int K>
__global__ void __launch_bounds__(VERIFY_LV_THREADS)
index_verify_lv_kernel(const __grid_constant__ IndexArgs a) {
  __shared__ unsigned long long s_q[VERIFY_LV_THREADS / 32][64];
  unsigned long long count = *a.seed_n;
  if (count > a.seed_cap) count = a.seed_cap;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  unsigned long long *q = s_q[warp];
  uint32_t nq = 0;   // warp-uniform
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + warp * 32u; i0 < count; i0 += stride) {
    const unsigned long long i = i0 + lane;
    bool pass = false;
    unsigned long long sd = 0;
    if (i + stride < count) {
      // the next seed's text words and accept masks towards L2 while this one is filtered (a seed's
      // window is 32 bytes somewhere in the packed text: HBM latency, and nothing else to hide it)
      const unsigned long long nx = a.seeds[i + stride];
      const int64_t s0n = (int64_t)(nx & ((1ULL << SEED_POS_BITS) - 1ULL)) - SEED_BIAS - K;
      const int64_t wn = s0n >> 4;
      if (wn >= 0 && (uint64_t)(wn + 3) < a.n16) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pk + wn));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pk + wn + 3));
      }
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.pacc + (uint32_t)(nx >> (SEED_POS_BITS + 3))));
    }
    if (i < count) {
      sd = a.seeds[i];
      LvCtx t;
      uint32_t slot;
      int pi;
      int64_t s_nom;
      lv_load_seed(a, sd, K, t, slot, pi, s_nom);
      pass = lv_prefilter<K>(t);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, pass);
    if (bal == 0u) continue;
    if (pass) q[nq + __popc(bal & lt)] = sd;
    nq += __popc(bal);
    if (nq >= 32u) {
      __syncwarp();
      const unsigned long long s2 = q[nq - 32u + lane];
      __syncwarp();
      nq -= 32u;
      lv_extend_seed<K>(a, true, s2);
    }
  }
  __syncwarp();
  lv_extend_seed<K>(a, lane < nq, lane < nq ? q[lane] : 0ULL);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.seed_count) atomicAdd(a.seed_count, count);
}

// level 1b for large pattern sets (substitutions): level 1 passes a sizeable fraction of the
// positions there; ask the L2-resident (key, distance) table before anything is queued.  The
// 16 loads of a lane are independent, so their latencies overlap.  Kept out of line: its 32
// temporaries must not cost the lookup loop its registers.
template <int L>
__device__ __noinline__ uint32_t pair_level1b_global(const IndexArgs &a, uint32_t h, int j, uint32_t c0,
                                                     uint32_t c1, uint32_t c2) {
  if (__all_sync(0xffffffffu, h == 0u)) return 0u;
  constexpr uint32_t L2 = 2 * L;
  constexpr uint32_t maskL = (1u << L2) - 1u;
  const uint32_t wsh2 = 32u - a.log_words2;
  const uint32_t salt2 = a.salt2[j];
  const uint32_t d2 = 2u * (uint32_t)(j + 1) * L;   // bit distance of the second piece
  uint32_t wrd[16], idx2[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t v0 = __funnelshift_r(c0, c1, 2 * i), v1 = __funnelshift_r(c1, c2, 2 * i), v2 = c2 >> (2 * i);
    const uint32_t hi = (d2 >= 32u) ? __funnelshift_r(v1, v2, d2) : __funnelshift_r(v0, v1, d2);
    const uint32_t key = (v0 & maskL) | ((hi & maskL) << L2);
    idx2[i] = key * a.mult2 + salt2;
    wrd[i] = ((h >> (15 - i)) & 1u) ? __ldg(a.bitmap2 + (idx2[i] >> wsh2)) : 0u;
  }
  uint32_t keep = 0u;
#pragma unroll
  for (int i = 0; i < 16; i++) keep |= ((wrd[i] >> (idx2[i] & 31u)) & 1u) << (15 - i);
  return keep;
}

// Substitution scan with compile-time piece length and distance count: the
// distances are L, 2L, .., ND*L, so the shifted windows G[t] = text >> 2t are
// shared between the lookups of different positions (position i at distance
// (j+1)L uses G[i + jL]): 16 + (ND-1)L funnel shifts per word instead of 16*ND.
// (Moving the word-index shift and the address LEA to the FMA pipe as IMAD.HI / IMAD
// was measured and is not worth it: the kernel sits at ~94 % of the shared-memory
// wavefront rate and ~87 % of the ALU rate at the same time, see DESIGN.md.)
// B1: a second kernel variant that also consults the shared-memory (key, distance) bitmap
// (level 1b) for the positions that passed level 1 -- worth it between ~24 k and 100 k
// entries (IUPAC primer sets, a few thousand primers); kept out of the B1 = false variant,
// whose lookup loop needs every register.
template <int L, int ND, bool B1>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
pair_scan_sub_kernel(const __grid_constant__ IndexArgs a) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  PairCta t;
  pair_cta_init(a, s_mem, t);
  const uint32_t wsh = 32u - a.log_words;
  constexpr uint32_t maskL = (1u << (2 * L)) - 1u;
  constexpr uint32_t mask2L = (L >= 8) ? 0xffffffffu : ((1u << (4 * L)) - 1u);
  constexpr uint32_t maskHi = mask2L & ~maskL;
  constexpr int NG = 16 + (ND - 1) * L;
  const uint32_t mult = a.mult;
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t salt[ND];
#pragma unroll
  for (int j = 0; j < ND; j++) salt[j] = a.salt[j];

  const int64_t c_step = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; c < t.n_chunks; c += c_step) {
    int64_t w;
    uint32_t cw[4], vm;
    pair_load_words(a, t, c, w, cw[0], cw[1], cw[2], vm);
    pair_prefetch_words(a, t, c + c_step);   // the next chunk's words on their way from HBM to L2 under the lookups
    cw[3] = 0u;
    uint32_t G[NG];
#pragma unroll
    for (int tt = 0; tt < NG; tt++) G[tt] = __funnelshift_r(cw[tt >> 4], cw[(tt >> 4) + 1], 2 * (tt & 15));
    uint32_t hm[ND];
#pragma unroll
    for (int j = 0; j < ND; j++) hm[j] = 0u;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t av = G[i] & maskL;
#pragma unroll
      for (int j = 0; j < ND; j++) {
        const uint32_t key = (j == 0) ? (G[i] & mask2L) : ((G[i + j * L] & maskHi) | av);
        hm[j] = pair_probe(t.bitmap, key, mult, salt[j], wsh, hm[j]);
      }
    }
    if (B1) {
#pragma unroll
      for (int j = 0; j < ND; j++) {
        const uint32_t h = hm[j] & vm;
        hm[j] = h ? pair_level1b(a, t, h, (uint32_t)j, cw[0], cw[1], cw[2]) : 0u;
      }
    } else if (a.b2_global) {
#pragma unroll
      for (int j = 0; j < ND; j++) hm[j] = pair_level1b_global<L>(a, hm[j] & vm, j, cw[0], cw[1], cw[2]);
    }
#pragma unroll
    for (int j = 0; j < ND; j++) pair_push<false>(a, t, w, (uint32_t)j, hm[j] & vm);
  }
  pair_cta_finish<false>(a, t);
}

template <int L, bool B1>
inline PairKernel pair_sub_kernel_nd(int nd) {
  switch (nd) {
    case 1: return pair_scan_sub_kernel<L, 1, B1>;
    case 2: return pair_scan_sub_kernel<L, 2, B1>;
    case 3: return pair_scan_sub_kernel<L, 3, B1>;
    case 4: return pair_scan_sub_kernel<L, 4, B1>;
  }
  return nullptr;
}
// specialised substitution kernel for (L, nd), or nullptr (the generic kernel serves);
// level1b: the variant that consults the shared-memory level 1b
inline PairKernel pair_sub_kernel(int L, int nd, bool level1b) {
  switch (L) {
    case 4: return level1b ? pair_sub_kernel_nd<4, true>(nd) : pair_sub_kernel_nd<4, false>(nd);
    case 5: return level1b ? pair_sub_kernel_nd<5, true>(nd) : pair_sub_kernel_nd<5, false>(nd);
    case 6: return level1b ? pair_sub_kernel_nd<6, true>(nd) : pair_sub_kernel_nd<6, false>(nd);
    case 7: return level1b ? pair_sub_kernel_nd<7, true>(nd) : pair_sub_kernel_nd<7, false>(nd);
    case 8: return level1b ? pair_sub_kernel_nd<8, true>(nd) : pair_sub_kernel_nd<8, false>(nd);
  }
  return nullptr;
}

// Edit verifier on the packed text: the recurrence of verify_seed (pm_scan.cuh) for one
// (slot, nominal window start) per thread.  The span t0 .. t_hi (at most m+3K+1 <= 42
// characters) is pulled into registers first, shifted so that character t0 sits at bit 0;
// the four per-code Shift-And masks live in shared memory (one row per thread), so a
// character costs one LDS instead of a chain of selects.
constexpr int VERIFY_IDX_THREADS = 128;
template <int K>
__global__ void __launch_bounds__(VERIFY_IDX_THREADS)
index_verify_kernel(const __grid_constant__ IndexArgs a) {
  __shared__ uint32_t s_u[5][VERIFY_IDX_THREADS];  // [code][thread]; row 4 = flagged character (0)
  unsigned long long count = *a.seed_n;
  if (count > a.seed_cap) count = a.seed_cap;
  s_u[4][threadIdx.x] = 0u;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long sd = a.seeds[i];
    const uint32_t slot = (uint32_t)(sd >> 40);
    const int64_t s_nominal = (int64_t)(sd & 0xffffffffffULL) - SEED_BIAS;
    const uint32_t m = a.plen[slot] & 0x7fu;
    const uint4 um = __ldg(reinterpret_cast<const uint4 *>(a.umask4) + slot);
    s_u[0][threadIdx.x] = um.x;
    s_u[1][threadIdx.x] = um.y;
    s_u[2][threadIdx.x] = um.z;
    s_u[3][threadIdx.x] = um.w;
    const uint32_t last = 1u << (m - 1);
    int64_t t_lo = s_nominal + (int64_t)m - 1 - K;
    int64_t t_hi = s_nominal + (int64_t)m - 1 + K;
    if (t_hi >= (int64_t)a.n) t_hi = (int64_t)a.n - 1;
    if (t_lo < 0) t_lo = 0;
    int64_t t0 = t_lo - (int64_t)(m + K);
    if (t0 < 0) t0 = 0;
    if (t_hi < t0) continue;
    const int span = (int)(t_hi - t0) + 1;     // <= m + 3K + 1 <= 42
    const int first_report = (int)(t_lo - t0);
    // characters t0 .. t0+47 as 3 x (codes, flags), character t0 at bit 0
    const int64_t wb = t0 >> 4;
    const uint32_t sh = 2u * (uint32_t)(t0 & 15);
    const uint2 p0 = pk_load(a, wb), p1 = pk_load(a, wb + 1), p2 = pk_load(a, wb + 2), p3 = pk_load(a, wb + 3);
    uint32_t cw[3] = {__funnelshift_r(p0.x, p1.x, sh), __funnelshift_r(p1.x, p2.x, sh), __funnelshift_r(p2.x, p3.x, sh)};
    uint32_t fw[3] = {__funnelshift_r(p0.y, p1.y, sh), __funnelshift_r(p1.y, p2.y, sh), __funnelshift_r(p2.y, p3.y, sh)};
    uint32_t M[K + 1];
#pragma unroll
    for (int l = 0; l <= K; l++) M[l] = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      uint32_t c = cw[j], f = fw[j];
      const int lim = min(16, span - 16 * j);
      for (int i = 0; i < lim; i++, c >>= 2, f >>= 2) {
        const uint32_t fl = f & 3u;
        const uint32_t u = s_u[fl ? 4u : (c & 3u)][threadIdx.x];
        // shift_and_inexact.cc:253-303, indels on; one word per pattern
        uint32_t old_prev = M[0];
        M[0] = ((M[0] << 1) | 1u) & u;
        if (!(fl & 2u)) {
#pragma unroll
          for (int l = 1; l <= K; l++) {
            const uint32_t old_l = M[l];
            uint32_t v = ((old_l << 1) | 1u) & u;
            v |= (old_prev << 1) | 1u | old_prev;  // substitution | insertion
            v |= (M[l - 1] << 1) | 1u | M[l - 1];  // deletion | lower level (new)
            M[l] = v;
            old_prev = old_l;
          }
        } else {
#pragma unroll
          for (int l = 1; l <= K; l++) M[l] = ((M[l] << 1) | 1u) & u;
        }
        if ((M[K] & last) && 16 * j + i >= first_report) index_emit(a, slot, t0 + 16 * j + i + 1);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.seed_count) atomicAdd(a.seed_count, count);
}

// stream bytes -> packed words.  lin[b]: bits 0-1 code, bit 2 bad, bit 3 eos,
// bit 4 ambiguous (a text byte that matches pattern symbols but is none of the four
// codes, e.g. a text 'R' under -w): packed as "bad" (the index kernels then never
// over-count a match through it) and its position is appended to `amb`, so that
// the windows around it can be re-examined on the raw text (amb_* kernels).
// A character that matches no pattern symbol gets a position-dependent code, so
// that a long run of them (the N blocks of a genome assembly) does not look like
// a homopolymer to level 1; the flags keep it from ever matching.
constexpr int PACK_WORDS_PER_THREAD = 8;
__global__ void __launch_bounds__(256)
pack_linear_kernel(const uint8_t *__restrict__ text, uint64_t n, const uint8_t *__restrict__ lin,
                   uint32_t eos_byte, uint2 *__restrict__ pk, uint64_t w_begin, uint64_t w_end,
                   unsigned long long *__restrict__ amb, unsigned long long *__restrict__ amb_n,
                   uint64_t amb_cap, int dna_fast) {
  __shared__ uint8_t s_lin[256];
  s_lin[threadIdx.x] = lin[threadIdx.x];
  __syncthreads();
  const uint64_t w_first = w_begin + (uint64_t)blockIdx.x * (256 * PACK_WORDS_PER_THREAD) + threadIdx.x;
#pragma unroll 2
  for (int r = 0; r < PACK_WORDS_PER_THREAD; r++) {
    const uint64_t w = w_first + (uint64_t)r * 256;
    if (w >= w_end) return;
    const uint64_t base = w * 16;
    uint32_t b[4];
    if (base + 16 <= n && (reinterpret_cast<uintptr_t>(text) & 15) == 0) {   // an adopted device pointer may be 4-byte aligned only
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + base));
      b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint32_t x = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint64_t i = base + 4 * j + q;
          x |= (uint32_t)((i < n) ? text[i] : (uint8_t)eos_byte) << (8 * q);
        }
        b[j] = x;
      }
    }
    if (dna_fast) {
      // A C G T only (the usual case): code = (byte >> 1) & 3, four characters per word.  The
      // bits outside the code must be 0x41 (A, C, G) or 0x50 (T); x * 0x01041040 gathers the
      // four 2-bit codes of a word (one per byte) into its top byte: byte j moves up by
      // 24 - 6j bits, and the other partial products stay below bit 24 or leave the word.
      uint32_t bad = 0, cw = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t x4 = b[j];
        const uint32_t is_t = (x4 >> 2) & ~(x4 >> 1) & 0x01010101u;
        bad |= (x4 & 0xF9F9F9F9u) ^ (0x41414141u + is_t * 15u);
        cw |= ((((x4 >> 1) & 0x03030303u) * 0x01041040u) >> 24) << (8 * j);
      }
      if (bad == 0) {
        pk[w] = make_uint2(cw, 0u);
        continue;
      }
    }
    uint32_t codes = 0, flags = 0, ambm = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t e = s_lin[(b[i >> 2] >> (8 * (i & 3))) & 0xffu];
      codes |= (e & 3u) << (2 * i);
      flags |= ((e >> 2) & 3u) << (2 * i);
      ambm |= (e >> 4) << i;
    }
    if (flags) {
      while (ambm) {
        const int i = __ffs(ambm) - 1;
        ambm &= ambm - 1;
        const unsigned long long at = atomicAdd(amb_n, 1ULL);
        if (at < amb_cap) amb[at] = base + (uint64_t)i;
      }
      const uint32_t bad = flags & 0x55555555u;
      const uint32_t noise = ((uint32_t)w * 0x9E3779B1u) ^ ((uint32_t)w >> 7) * 0x85EBCA6Bu;
      codes = (codes & ~(bad * 3u)) | (noise & (bad * 3u));
    }
    // a position past the end of the stream is an end-of-sequence character
    pk[w] = make_uint2(codes, flags);
  }
}

// ---------------------------------------------------------------------------
// Ambiguous text characters (pack_linear_kernel's list): every window / alignment
// that contains one is decided on the raw text with the general class masks of
// pm_scan.cuh, for all patterns.  Rare by assumption (a handful in a genome
// assembly); the host falls back to the bit-plane kernels when the list overflows.
//
// substitutions: every window start x-m+1 .. x of every slot
__global__ void __launch_bounds__(256)
amb_sub_kernel(const ScanArgs a, const unsigned long long *__restrict__ amb,
               const unsigned long long *__restrict__ n_amb_ptr, uint64_t amb_cap, uint32_t k) {
  // the count is read on the device (no host round trip); a list that overflowed is not worth
  // walking, the host falls back to the bit-plane kernels
  uint64_t n_amb = *n_amb_ptr;
  if (n_amb > amb_cap) return;
  for (uint64_t ai = blockIdx.x; ai < n_amb; ai += gridDim.x) {
    const int64_t x = (int64_t)amb[ai];
    for (uint32_t idx = threadIdx.x; idx < a.n_slots * 32u; idx += blockDim.x) {
      const uint32_t slot = idx >> 5, o = idx & 31u;
      const uint32_t m = a.plen[slot] & 0x7fu;
      if (o >= m) continue;
      const int64_t s = x - (int64_t)o;
      if (s < 0 || (uint64_t)(s + m) > a.n) continue;
      const uint64_t *U = a.umask + (size_t)slot * a.n_cls;
      uint32_t mis = 0;
      bool ok = true;
      for (uint32_t j = 0; j < m; j++) {
        const uint32_t ch = a.text[s + j];
        if (ch == a.eos) { ok = false; break; }
        if (!((__ldg(U + a.cls[ch]) >> j) & 1ull) && ++mis > k) { ok = false; break; }
      }
      if (ok) emit_candidate(a, slot, s + m);
    }
  }
}

// edits: seeds for verify_seed covering every alignment end x .. x+m+K-1 of every slot
__global__ void __launch_bounds__(256)
amb_seed_kernel(const ScanArgs a, const unsigned long long *__restrict__ amb, uint64_t amb_begin,
                uint64_t amb_end, int K) {
  for (uint64_t ai = amb_begin + blockIdx.x; ai < amb_end; ai += gridDim.x) {
    const int64_t x = (int64_t)amb[ai];
    for (uint32_t slot = threadIdx.x; slot < a.n_slots; slot += blockDim.x) {
      const int64_t m = a.plen[slot] & 0x7fu;
      for (int64_t tc = x + K; tc - K <= x + m + K - 1; tc += 2 * K + 1) {
        int64_t s_nominal = tc - m + 1;
        if (s_nominal < 0) s_nominal = 0;
        const unsigned long long at = atomicAdd(a.seed_n, 1ULL);
        if (at < a.seed_cap) a.seeds[at] = ((unsigned long long)slot << 40) | (unsigned long long)(s_nominal + SEED_BIAS);
        else *a.seed_overflow = 1ULL;
      }
    }
  }
}

}  // namespace pm
