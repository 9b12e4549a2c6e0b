// pm_index.cuh -- text-driven q-gram index scan (sm_100a).
//
// The bit-plane kernels of pm_scan.cuh touch every (pattern, position) pair.
// For DNA-like alphabets that is far more work than the problem needs: with at
// most k errors one of the k+1 pieces of a pattern occurs exactly, so a text
// position only has to look at the patterns whose piece starts with the q-gram
// found there.  This is the role the reference gives its keyword tree / hash
// table front ends (exact_bases.cc, exact_halves.cc, hash_table.cc); here it
// is one flat table:
//
//   bucket[code]        = first entry of the q-gram `code` (4^q + 1 uint32)
//   entry               = slot << 6 | offset of the piece in its pattern
//   q                   = min(shortest piece, 12), chosen on the host
//
// and the text is kept 2 bits per character (+ 2 flag bits: "matches no
// pattern symbol" and "end of sequence"), 16 characters per 8-byte word, so a
// verification is a few loads, an XOR and a population count.
//
// Exactness.  The candidate set is the one of pm_scan.cuh / the reference:
//   substitutions: the verifier counts the mismatches of the whole window and
//     rejects windows holding an end-of-sequence character -- the level-k hit
//     condition of shift_and_inexact.cc:253-326 without indels;
//   edits: the table only proposes seeds (exact occurrence of a whole piece);
//     the reference's own recurrence decides (verify_seed_packed below).
// Pigeonhole makes the seeds a superset in both cases.
//
// Applicable when every text byte matches at most one pattern symbol, there
// are at most 4 symbols and no pattern is longer than 32 (host checks); other
// inputs (IUPAC patterns, text wildcards, long patterns) use the bit-plane
// kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pm_scan.cuh"

namespace pm {

constexpr int IDX_THREADS = 128;
constexpr int IDX_SEG = 128;                        // q-gram end positions per thread
constexpr int IDX_BLOCK_POS = IDX_THREADS * IDX_SEG;  // 16384 positions per block
constexpr int IDX_QMAX = 12;
constexpr int IDX_SMEM_Q = 7;                       // bucket array in shared memory up to this q

struct IndexArgs {
  const uint2 *pk;          // [n16] x = 2-bit codes of 16 characters, y = flags (bit0 bad, bit1 eos)
  uint64_t n;               // shard length (characters)
  uint64_t n16;             // words in pk
  const uint32_t *bucket;   // [4^q + 1]
  const uint32_t *entries;  // slot << 6 | piece offset
  const uint64_t *pat2;     // [n_slots] 2-bit codes of the pattern
  const uint8_t *plen;      // [n_slots]
  const uint64_t *piece_end;  // [n_slots] bit i: character i ends a piece
  const uint32_t *umask4;   // [n_slots][4] Shift-And masks per text code (edit verifier)
  unsigned long long *out;  // candidates: slot << 40 | pos
  unsigned long long *count;
  uint64_t cap;
  int64_t lo, hi;           // emit candidates with lo <= pos < hi
  int64_t p_begin, p_end;   // q-gram END positions handled by this launch
  uint32_t q;
  uint32_t k;
  // edit mode: seeds instead of candidates
  unsigned long long *seeds, *seed_n, *seed_overflow, *seed_count;
  uint64_t seed_cap;
  unsigned long long *work_count;  // table entries examined
};

// 16-character words of the packed text.  Characters past the end read as end
// of sequence (the pack kernel fills whole tiles).
__device__ __forceinline__ uint2 pk_load(const IndexArgs &a, int64_t w) {
  if (w < 0 || (uint64_t)w >= a.n16) return make_uint2(0u, 0xffffffffu);
  return __ldg(a.pk + w);
}

// codes and flags of the `len` (<= 32) characters starting at s, right-aligned
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

// EDIT = false: substitutions-only verifier inline, candidates out.
// EDIT = true : whole-piece check, seeds out (verified by index_verify_kernel);
//               seeds are queued per warp in shared memory and flushed with one
//               global atomic per flush.
constexpr int IDX_QCAP = 256;  // per-warp seed queue entries

template <bool EDIT>
__global__ void __launch_bounds__(IDX_THREADS)
qgram_scan_kernel(const IndexArgs a) {
  extern __shared__ uint32_t s_bucket[];  // 4^q + 1 entries when q <= IDX_SMEM_Q
  __shared__ unsigned long long s_queue[EDIT ? IDX_THREADS / 32 : 1][EDIT ? IDX_QCAP : 1];
  __shared__ uint32_t s_qn[IDX_THREADS / 32];
  const uint32_t q = a.q;
  const uint32_t n_buckets = 1u << (2 * q);
  const bool in_smem = q <= (uint32_t)IDX_SMEM_Q;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_qn[warp] = 0;
  if (in_smem)
    for (uint32_t i = threadIdx.x; i <= n_buckets; i += blockDim.x) s_bucket[i] = a.bucket[i];
  __syncthreads();
  const uint32_t *bucket = in_smem ? s_bucket : a.bucket;
  const uint32_t qmask = n_buckets - 1u;

  // this thread's q-gram end positions [p0, p1) (empty past the end of the launch)
  const int64_t p0 = a.p_begin + ((int64_t)blockIdx.x * IDX_THREADS + threadIdx.x) * IDX_SEG;
  int64_t p1 = p0 + IDX_SEG;
  if (p1 > a.p_end) p1 = a.p_end;
  // start one word early: warms the rolling code with >= 16 >= q-1 characters
  const int64_t w_first = (p0 >> 4) - 1;
  uint32_t code = 0, run = 0;
  unsigned long long n_entries = 0;

  auto flush = [&]() {
    if (!EDIT) return;
    __syncwarp();
    const uint32_t qn = min(s_qn[warp], (uint32_t)IDX_QCAP);
    unsigned long long base = 0;
    if (lane == 0 && qn) base = atomicAdd(a.seed_n, (unsigned long long)qn);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t i = lane; i < qn; i += 32) {
      if (base + i < a.seed_cap) a.seeds[base + i] = s_queue[EDIT ? warp : 0][i];
      else *a.seed_overflow = 1ULL;
    }
    __syncwarp();
    if (lane == 0) s_qn[warp] = 0;
    __syncwarp();
  };

  // every thread of the warp walks the same number of words: the flush points are uniform
  constexpr int NW = IDX_SEG / 16 + 2;
  for (int wi = 0; wi < NW; ++wi) {
    const int64_t w = w_first + wi;
    const uint2 v = pk_load(a, w);
    uint32_t cw = v.x, fw = v.y;
    const int64_t base = w << 4;
#pragma unroll 4
    for (int i = 0; i < 16; i++, cw >>= 2, fw >>= 2) {
      code = ((code << 2) | (cw & 3u)) & qmask;
      run = (fw & 3u) ? 0u : run + 1u;
      const int64_t p = base + i;  // the q-gram ends here
      if (run < q || p < p0 || p >= p1) continue;
      uint32_t e0 = bucket[code];
      const uint32_t e1 = bucket[code + 1];
      for (; e0 < e1; ++e0) {
        const uint32_t ent = __ldg(a.entries + e0);
        const uint32_t slot = ent >> 6, off = ent & 63u;
        const uint32_t m = a.plen[slot] & 0x7fu;
        const int64_t s = p - (int64_t)q + 1 - (int64_t)off;  // window start
        n_entries++;
        // substitutions: the window must lie inside the stream; edits: a nominal window may
        // hang over the end (trailing deletions), the verifier clamps
        if (s < 0 || (!EDIT && (uint64_t)(s + m) > a.n)) continue;
        uint64_t codes, flags;
        pk_window(a, s, codes, flags);
        const uint64_t lenmask = (m >= 32) ? ~0ULL : ((1ULL << (2 * m)) - 1ULL);
        const uint64_t diff = codes ^ __ldg(a.pat2 + slot);
        const uint64_t mis = ((diff | (diff >> 1)) | flags) & 0x5555555555555555ULL & lenmask;
        if (!EDIT) {
          if (__popcll(mis) <= (int)a.k && (flags & 0xAAAAAAAAAAAAAAAAULL & lenmask) == 0)
            index_emit(a, slot, s + m);
        } else {
          // the whole piece must match: characters off .. piece end
          const uint64_t pe = __ldg(a.piece_end + slot) >> off;
          const uint32_t plen_piece = (uint32_t)__ffsll((long long)pe);  // 1-based: piece length
          const uint64_t pm = ((plen_piece >= 32) ? ~0ULL : ((1ULL << (2 * plen_piece)) - 1ULL)) << (2 * off);
          if ((mis & pm) == 0) {
            const unsigned long long sd = ((unsigned long long)slot << 40) | (unsigned long long)s;
            const uint32_t at = atomicAdd(&s_qn[warp], 1u);
            if (at < (uint32_t)IDX_QCAP) {
              s_queue[EDIT ? warp : 0][at] = sd;
            } else {  // queue full: straight to the list
              const unsigned long long g = atomicAdd(a.seed_n, 1ULL);
              if (g < a.seed_cap) a.seeds[g] = sd; else *a.seed_overflow = 1ULL;
            }
          }
        }
      }
    }
    if (EDIT) {
      __syncwarp();
      if (s_qn[warp] >= (uint32_t)(IDX_QCAP / 2)) flush();
    }
  }
  flush();
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) n_entries += __shfl_xor_sync(0xffffffffu, n_entries, d);
  if (lane == 0 && a.work_count) atomicAdd(a.work_count, n_entries);
}

// Edit verifier on the packed text: the recurrence of verify_seed (pm_scan.cuh)
// with the character masks of the four codes in registers.
template <int K>
__global__ void __launch_bounds__(128)
index_verify_kernel(const IndexArgs a) {
  unsigned long long count = *a.seed_n;
  if (count > a.seed_cap) count = a.seed_cap;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long sd = a.seeds[i];
    const uint32_t slot = (uint32_t)(sd >> 40);
    const int64_t s_nominal = (int64_t)(sd & 0xffffffffffULL);
    const uint32_t m = a.plen[slot] & 0x7fu;
    const uint4 um = __ldg(reinterpret_cast<const uint4 *>(a.umask4) + slot);
    const uint32_t U[4] = {um.x, um.y, um.z, um.w};
    const uint32_t last = 1u << (m - 1);
    int64_t t_lo = s_nominal + (int64_t)m - 1 - K;
    int64_t t_hi = s_nominal + (int64_t)m - 1 + K;
    if (t_hi >= (int64_t)a.n) t_hi = (int64_t)a.n - 1;
    if (t_lo < 0) t_lo = 0;
    int64_t t0 = t_lo - (int64_t)(m + K);
    if (t0 < 0) t0 = 0;
    uint32_t M[K + 1];
#pragma unroll
    for (int l = 0; l <= K; l++) M[l] = 0;
    // up to 32+3K+1+32 < 80 characters: five packed words cover any alignment
    const int64_t wb = t0 >> 4;
    uint2 pw[6];
#pragma unroll
    for (int j = 0; j < 6; j++) pw[j] = pk_load(a, wb + j);
#pragma unroll
    for (int j = 0; j < 6; j++) {
      uint32_t cw = pw[j].x, fw = pw[j].y;
      const int64_t base = (wb + j) << 4;
      if (base > t_hi) break;
#pragma unroll 4
      for (int i = 0; i < 16; i++, cw >>= 2, fw >>= 2) {
        const int64_t t = base + i;
        if (t < t0 || t > t_hi) continue;
        const uint32_t fl = fw & 3u;
        const uint32_t c = cw & 3u;
        const uint32_t u = fl ? 0u : (c == 0 ? U[0] : c == 1 ? U[1] : c == 2 ? U[2] : U[3]);
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
        if (t >= t_lo && (M[K] & last)) index_emit(a, slot, t + 1);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.seed_count) atomicAdd(a.seed_count, count);
}

// stream bytes -> packed words.  lin[b]: bits 0-1 code, bit 2 bad, bit 3 eos.
__global__ void __launch_bounds__(256)
pack_linear_kernel(const uint8_t *__restrict__ text, uint64_t n, const uint8_t *__restrict__ lin,
                   uint32_t eos_byte, uint2 *__restrict__ pk, uint64_t w_begin, uint64_t w_end) {
  __shared__ uint8_t s_lin[256];
  s_lin[threadIdx.x] = lin[threadIdx.x];
  __syncthreads();
  const uint64_t w = w_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= w_end) return;
  const uint64_t base = w * 16;
  uint32_t b[4];
  if (base + 16 <= n) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + base));
    b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t x = 0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const uint64_t i = base + 4 * j + r;
        x |= (uint32_t)((i < n) ? text[i] : (uint8_t)eos_byte) << (8 * r);
      }
      b[j] = x;
    }
  }
  uint32_t codes = 0, flags = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint32_t e = s_lin[(b[i >> 2] >> (8 * (i & 3))) & 0xffu];
    codes |= (e & 3u) << (2 * i);
    flags |= ((e >> 2) & 3u) << (2 * i);
  }
  // a position past the end of the stream is an end-of-sequence character
  pk[w] = make_uint2(codes, flags);
}

}  // namespace pm
