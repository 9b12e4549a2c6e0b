// pm_b200.cu -- context, host logic and the extern "C" boundary (include/pm_b200.h)
// of the B200-native PrimerMatch database scan.  Kernels: pm_index.cuh (pack + pair-seed
// index scan), pm_fasta.cuh (FASTA normalisation), pm_planes.cuh / pm_scan.cuh (bit-plane
// scan, .sqz unpack), pm_halves.cuh (exact_halves), pm_align.cuh (chains + alignment).
//
// There is no CPU path in here: every entry point that touches text or
// patterns runs on the CUDA device of the context or fails.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <algorithm>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

#include "../../include/pm_b200.h"
#include "pm_align.cuh"
#include "pm_planes.cuh"
#include "pm_scan.cuh"
#include "pm_index.cuh"
#include "pm_fasta.cuh"
#include "pm_halves.cuh"
#include "pm_pcr.hpp"
#include "pm_reader.hpp"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define PM_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(PM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));           \
  } while (0)

// ---------------------------------------------------------------------------
// IUPAC compatibility, util.cc:121-185: rows[w] lists the text characters a
// pattern character w is compatible with.  The table is the reference's, not
// the textbook one (the N row lacks B, there is an X row).
struct IupacRow { char w; const char *members; };
const IupacRow kIupac[] = {
    {'A', "ARMWDHVN"},  {'B', "GTUCYKSBN"}, {'C', "CYMSBHVN"}, {'D', "GATURWKDN"},
    {'G', "GRKSBDVN"},  {'H', "ACTUMYWHN"}, {'K', "GTKBDN"},   {'M', "ACMHVN"},
    {'N', "ACGTURYKMSWVDHVN"},              {'R', "GARDVN"},   {'S', "GCSBVN"},
    {'T', "TUYKWVDHN"}, {'U', "UTYKWVDHN"}, {'V', "GCARSMVN"}, {'W', "ATUWDHN"},
    {'Y', "TUCYBHN"},   {'X', "MRWSYKVHDBXN"}};

struct Compat {
  bool has_row[128];
  bool ok[128][128];
  Compat() {
    memset(has_row, 0, sizeof has_row);
    memset(ok, 0, sizeof ok);
    for (const IupacRow &r : kIupac) {
      const int up = r.w, lo = r.w - 'A' + 'a';
      has_row[up] = has_row[lo] = true;
      for (const char *p = r.members; *p; ++p) {
        ok[up][(int)*p] = true;
        ok[lo][*p - 'A' + 'a'] = true;
      }
    }
  }
};
const Compat &compat() {
  static Compat c;
  return c;
}

// The scanner's character relation: shift_and_inexact.cc:126-143 and
// shift_and.cc:107-119.  With wildcards a pattern character that has an IUPAC
// row matches the members of the row (a text 'N' only under -W); everything
// else matches itself only.
bool scan_match(unsigned pc, unsigned tc, bool wc, bool tn) {
  const Compat &c = compat();
  if (wc && pc < 128 && c.has_row[pc]) {
    if (tc >= 128 || !c.ok[pc][tc]) return false;
    if (tc == 'N' && !tn) return false;
    return true;
  }
  return pc == tc;
}

unsigned char revcomp_char(unsigned char ch) {  // util.cc:319-348
  static const char *pairs = "ATCGGCTAUAMKRYWWSSYRKMVBHDDHBV";
  for (const char *p = pairs; *p; p += 2) {
    if (ch == (unsigned char)p[0]) return (unsigned char)p[1];
    if (ch == (unsigned char)(p[0] - 'A' + 'a')) return (unsigned char)(p[1] - 'A' + 'a');
  }
  return ch;
}

// ---------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return PM_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(PM_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    cap = want;
    return PM_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
// a buffer that lives for one call (freed on every way out of the function)
struct LocalBuf : DevBuf {
  ~LocalBuf() { release(); }
};

struct Pattern {
  std::string s;
  uint32_t id;
  int esb, eeb;
};

constexpr size_t kStageBytes = 64u << 20;       // pinned staging buffer (x2)
constexpr uint64_t kDirectOutBytes = 64ull << 20;  // records written straight to page-locked memory up to this bound
constexpr size_t kPoolBytes = 32u << 20;        // initial pool of the long-chain alignments
constexpr int64_t kRightGuard = 1 << 16;
constexpr uint64_t kAmbCap = 1u << 20;          // ambiguous text characters the index path re-examines
constexpr int kGuardTooShort = -101;           // internal: an owned chain runs into the end of the right guard
constexpr int kFallbackToPlanes = -100;        // internal: too many of them, rescan with the bit-plane kernels

}  // namespace

struct pm_ctx {
  pm_options opt;
  int index = 0;  // resolved PM_INDEX_*
  cudaStream_t stream = 0;
  cudaStream_t copy_stream = 0;   // host -> device copies only, back to back
  cudaStream_t prep_stream = 0;   // unpack / pack kernels of the chunks that have arrived
  std::vector<Pattern> pats;
  bool inited = false;
  bool index_auto = false;   // pm_create got PM_INDEX_AUTO with k > 0: resolved in pm_init
  pm_ctx *inner = nullptr;   // exact_halves / exact_bases: the exact index over the pattern parts
  bool inner_bases = false;  // ... exact_bases
  bool unordered_ok = false; // this IS such an inner context: run_scan may leave the candidates unsorted
  int align_rules = PM_ALIGN_PRIMER_MATCH;  // whose alignment an exact (k == 0) hit gets, pm_set_alignment_rules
  int alphabet_size = 256;   // CharacterProducer::size() of the database, pm_set_alphabet_size
  DevBuf d_half_rec, d_half_key, d_half_key2, d_half_ord, d_half_ord2;
  int max_len = 0;
  // symbols
  int n_sym = 0;
  int n_cls = 0;
  // device tables
  DevBuf d_symmask, d_poff, d_plen, d_patchars, d_lconst, d_rconst, d_compat, d_cls, d_umask,
      d_piece, d_slotid;
  // text
  DevBuf d_text_own, d_planes, d_sqz, d_sqztab;
  // device-side FASTA normalisation (pm_fasta.cuh)
  DevBuf d_fa_raw, d_fa_map, d_fa_scan, d_fa_cnt, d_fa_chdr, d_fa_start, d_fa_nout, d_fa_nhdr, d_fa_ooff, d_fa_hoff,
      d_fa_key, d_fa_hs, d_fa_he, d_fa_tmp;
  std::vector<int64_t> seq_key, seq_hdr_off, seq_hdr_len;  // sequence table of the last FASTA text
  std::string hdr_blob;                 // header lines (gathered from the device, or <db>.hdr)
  std::vector<unsigned long long> hdr_blob_off;
  bool hdr_ready = false, hdr_from_file = false;
  uint64_t last_uniq = 0;               // candidates of the last pm_find_patterns (d_keys), for pm_last_candidates
  std::vector<pm_candidate> h_cands;
  int db_kind = 0;                      // pm_set_text_database: 1 FASTA, 2 .seq, 3 .sqn, 4 .sqz
  pm::ChunkReader *reader = nullptr;    // ... and its reader (kept until the next text: its buffers are DMA sources)
  const uint8_t *d_text = nullptr;
  uint64_t n = 0;
  int64_t offset = 0;
  uint32_t n_tiles = 0;
  void *stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  // planes become ready chunk by chunk: tiles [previous tile_end, tile_end) after `ev`
  // (ev == nullptr: packed on the main stream, no wait needed)
  struct Chunk { uint32_t tile_end; cudaEvent_t ev; };
  std::vector<Chunk> chunks;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  bool pack_timed = false;
  // work buffers
  DevBuf d_pool;   // storage of the chain alignments whose band exceeds ALIGN_BAND_MAX (bump-allocated)
  bool out_direct = false;   // the last call wrote its records straight to h_out
  int n_sm = 148;
  DevBuf d_keys, d_keys2, d_head, d_incl, d_chains, d_raw, d_hits, d_hits2,
      d_sortk, d_sortk2, d_idx, d_idx2, d_tmp, d_cands, d_seeds;
  double seed_rate = 0;  // expected piece occurrences per base on uniform ACGT text
  double idx_seed_rate = 0;  // same for the pair index (pair occurrences)
  uint64_t idx_entries = 0;
  // pair-seed index path (pm_index.cuh)
  int scan_path = PM_SCAN_AUTO;  // requested
  bool use_index = false;        // the loaded text is packed for / scanned by the pair index
  bool index_ok = false;         // resolved in pm_init: the patterns can use the pair index
  bool idx_has_amb = false;      // some text byte would be ambiguous (wildcards): pack lists them
  bool idx_dna_fast = false;     // text codes are (byte >> 1) & 3 for A, C, G, T: table-free pack
  bool idx_edit_special = false; // edits: pair_scan_edit_kernel<L, K> + index_verify_lv_kernel<K> serve (else the run-time pair)
  DevBuf d_amb;
  int idx_q = 0;                 // piece length L of the pair index
  DevBuf d_pk, d_bucket, d_entries, d_pacc, d_umask4, d_lin, d_bitmap, d_bitmap2;
  pm::IndexArgs idx_proto;       // level-1 / level-2 parameters filled by pm_init
  unsigned long long *d_counters = nullptr;  // [16]; [8] = ambiguous text characters of the packed text
  unsigned long long *h_counters = nullptr;  // pinned [16]
  void *h_out = nullptr;                     // pinned result buffer
  size_t h_out_cap = 0;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  pm_stats stats;
};

namespace {

using namespace pm;

int ensure_host_out(pm_ctx *c, size_t bytes) {
  if (bytes <= c->h_out_cap) return PM_OK;
  if (c->h_out) cudaFreeHost(c->h_out);
  c->h_out = nullptr;
  c->h_out_cap = 0;
  const size_t want = bytes + bytes / 2 + 4096;
  PM_CUDA(cudaMallocHost(&c->h_out, want));
  c->h_out_cap = want;
  return PM_OK;
}

AlignEnv make_env(pm_ctx *c) {
  AlignEnv e;
  e.text = c->d_text;
  e.n = (int64_t)c->n;
  e.pat_chars = c->d_patchars.as<uint8_t>();
  e.plen = c->d_plen.as<uint8_t>();
  e.lconst = c->d_lconst.as<int16_t>();
  e.rconst = c->d_rconst.as<int16_t>();
  e.compat = c->d_compat.as<uint32_t>();
  e.k = c->opt.k;
  e.indels = c->opt.indels;
  e.wc = c->opt.wildcards;
  e.tn = c->opt.text_n;
  e.eos = c->opt.eos;
  return e;
}

// Device-side counters of a context (c->d_counters / c->h_counters): the tail of a call reads
// its counts from here instead of from the host, so that one call costs one read-back.
enum : int {
  CTR_COUNT = 0,      // raw candidates appended by the scan kernels
  CTR_SEEDS = 1,      // edit mode: seeds verified (all batches)
  CTR_UNIQ = 2,       // sorted unique candidates
  CTR_BOGUS = 3,
  CTR_RAW = 4,        // hits after the chain alignment / exact_halves rule
  CTR_WORK = 5,
  CTR_SEED_N = 6,     // seeds of the current batch
  CTR_SEED_OVF = 7,
  CTR_AMB = 8,        // ambiguous text characters of the packed text
  CTR_CUT = 9,        // an owned chain reaches the end of the right guard
  CTR_HALF_OK = 10,   // exact_halves: occurrences that extended
  CTR_FA = 11,        // [11..14] FASTA normalisation carry
  CTR_CHAINS = 16,
  CTR_POOL = 17,      // bytes of the long-chain pool asked for
  CTR_NHITS = 18,     // hits written by the last tail
  CTR_N = 32
};

__global__ void store_u64_kernel(unsigned long long *p, unsigned long long v) { *p = v; }

// final alignment of one reported hit: primer_match.cc:1132-1153.  R = 0, so the band is
// 2k+1 wide and the whole alignment lives in per-thread memory.  Threads past the hit count
// pad the sort keys (the sort runs over `bound` items, an upper bound known to the host).
__global__ void __launch_bounds__(64)
final_align_kernel(const AlignEnv env, const RawHit *raw, const unsigned long long *n_ptr, uint64_t bound,
                   const uint32_t *slot_id, int64_t offset, pm_hit *hits, unsigned long long *sortkey,
                   unsigned long long *n_bogus, unsigned long long *n_hits, int use_ord) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound) return;
  const uint64_t n = *n_ptr;
  if (i == 0) *n_hits = n;
  if (i >= n) {
    sortkey[i] = ~0ULL;
    return;
  }
  const RawHit h = raw[i];
  AlignResult r;
  pm_hit out;
  memset(&out, 0, sizeof out);
  uint8_t best[ALIGN_LOCAL_BEST];
  int row_a[ALIGN_BAND_MAX + 2], row_b[ALIGN_BAND_MAX + 2];
  editdist_align_band(env, h.slot, h.end, h.end, best, row_a, row_b, out.codes, PM_MAX_CODES, r);
  out.start = r.start + offset;
  out.end = r.end + offset;
  out.first = h.first + offset;
  out.pattern_id = slot_id[h.slot];
  out.editdist = r.editdist;
  out.value = r.val;
  out.ncodes = (uint32_t)(r.ncodes < PM_MAX_CODES ? r.ncodes : PM_MAX_CODES);
  // pattern_alignment stores the codes in traceback order reversed into
  // pattern order (pattern_alignment.cc:592-600): the DP walks right-to-left,
  // so traceback order already is left-to-right.
  if (r.editdist < 0 || r.editdist > env.k) atomicAdd(n_bogus, 1ULL);
  hits[i] = out;
  // order: (first, pattern id); first < 2^40, id < 2^24 -- the order filter_bitvec reports its chains
  // in, up to hits that begin at one position.  exact_halves / exact_bases: the order of the
  // occurrences they extend (RawHit::ord).
  sortkey[i] = use_ord ? h.ord
                       : (((unsigned long long)(out.first & 0xffffffffffLL) << 24) |
                          (unsigned long long)(out.pattern_id & 0xffffffu));
}

// exact indexes (k == 0): every candidate is a hit.  exact_alignment
// (pattern_alignment.cc:40-68) / exact_wc_alignment (pattern_alignment.cc:70-93).
__global__ void exact_hits_kernel(const AlignEnv env, const unsigned long long *keys, const unsigned long long *n_ptr,
                                  uint64_t bound, const uint32_t *slot_id, int64_t offset, pm_hit *hits,
                                  unsigned long long *sortkey, unsigned long long *n_bogus,
                                  unsigned long long *n_hits, int pcr_rules) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bound) return;
  const uint64_t n = *n_ptr;
  if (i == 0) *n_hits = n;
  if (i >= n) {
    sortkey[i] = ~0ULL;
    return;
  }
  const unsigned long long key = keys[i];
  const uint32_t slot = (uint32_t)(key >> 40);
  const int64_t pos = (int64_t)(key & 0xffffffffffULL);
  const int m = env.plen[slot] & 0x7f;
  const uint8_t *pat = env.pat_chars + (size_t)slot * PAT_STRIDE;
  pm_hit out;
  memset(&out, 0, sizeof out);
  int sub = 0;
  for (int j = 0; j < m; j++) {
    uint8_t code = AC_EQUAL;
    if (env.wc) {
      const uint32_t tc = env.text[pos - m + j], pc = pat[j];
      // exact_wc_alignment asks iupac_compatible(text, pattern) (pattern_alignment.cc:79), the
      // editdist_alignment pcr_match makes at k = 0 too asks iupac_compatible(pattern, text)
      // (pattern_alignment.cc:317); without an error to spend, that is its whole difference
      const bool compat = pcr_rules ? compat_bit(env.compat, pc, tc) : compat_bit(env.compat, tc, pc);
      if (tc == pc) code = AC_EQUAL;
      else if (compat && (env.tn || tc != 'N')) code = AC_WC_EQUAL;
      else { code = AC_SUBST; sub++; }
    }
    if (j < PM_MAX_CODES) out.codes[j] = code;
  }
  out.start = pos - m + offset;
  out.end = pos + offset;
  out.first = pos + offset;
  out.pattern_id = slot_id[slot];
  out.editdist = sub;
  out.value = 0;
  out.ncodes = (uint32_t)m;
  if (sub > 0) atomicAdd(n_bogus, 1ULL);
  hits[i] = out;
  sortkey[i] = ((unsigned long long)(out.end & 0xffffffffffLL) << 24) |
               (unsigned long long)(out.pattern_id & 0xffffffu);
}

// hits in sorted order, written 8 bytes per thread so that a warp writes whole lines: `out`
// may be page-locked host memory (the records then cross PCIe once, without a count on the host)
static_assert(sizeof(pm_hit) % 8 == 0, "pm_hit is copied in 8-byte words");
constexpr uint32_t HIT_WORDS = sizeof(pm_hit) / 8;
__global__ void gather_hits_kernel(const pm_hit *in, const uint32_t *idx, const unsigned long long *n_ptr,
                                   pm_hit *out) {
  const uint64_t n_words = *n_ptr * HIT_WORDS;
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(in);
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(out);
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_words; f += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = f / HIT_WORDS;
    const uint32_t j = (uint32_t)(f - i * HIT_WORDS);
    dst[f] = src[(uint64_t)idx[i] * HIT_WORDS + j];
  }
}

__global__ void iota_kernel(uint32_t *p, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t)i;
}

__global__ void keys_to_cands_kernel(const unsigned long long *keys, uint64_t n,
                                     const uint32_t *slot_id, int64_t offset, pm_candidate *out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  pm_candidate c;
  c.pos = (int64_t)(key & 0xffffffffffULL) + offset;
  c.pattern_id = slot_id[(uint32_t)(key >> 40)];
  c.reserved = 0;
  out[i] = c;
}

inline unsigned nblocks(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int slot_bits(size_t n_slots) {
  int b = 1;
  while ((1ull << b) < n_slots) b++;
  return b;
}

// ---------------------------------------------------------------------------
// dense scan over local candidate positions [lo, hi): fills c->d_keys (sorted, unique).  The
// number of candidates is left in d_counters[CTR_UNIQ]; *n_out is that number when exact_count
// is set (one more read-back), otherwise an upper bound (the candidates before duplicates were
// removed), which is all the tail of pm_find_patterns needs on the host.  One host
// synchronisation: the scan's counters (candidates, seed-list / candidate-buffer overflow).
int run_scan(pm_ctx *c, int64_t lo, int64_t hi, uint64_t *n_out, bool exact_count) {
  *n_out = 0;
  c->last_uniq = 0;
  PM_CUDA(cudaMemsetAsync(c->d_counters + CTR_UNIQ, 0, sizeof(unsigned long long), c->stream));
  if (c->n == 0 || hi <= lo || c->pats.empty()) return PM_OK;
  const size_t n_slots = c->pats.size();
  const int K = c->opt.k;
  const bool edit = (K > 0 && c->opt.indels);
  // window starts that can end in [lo-1, hi-1)
  int64_t q_lo = lo - 1 - (int64_t)c->max_len - K;
  if (q_lo < 0) q_lo = 0;
  int64_t q_hi = hi - 1 + K;  // exclusive bound on window starts (edit: nominal start + K)
  if (q_hi > (int64_t)c->n) q_hi = (int64_t)c->n;
  if (q_hi <= q_lo) return PM_OK;
  const uint32_t tile0 = (uint32_t)(q_lo / TILE_POS);
  const uint32_t tile1 = (uint32_t)((q_hi - 1) / TILE_POS) + 1;
  const uint32_t nt = tile1 - tile0;

  // pattern chunks: enough blocks to fill the chip several times over
  int dev = c->opt.device, n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t want_blocks = (uint64_t)n_sm * 32;
  uint32_t chunks = 1;
  if (nt < want_blocks) chunks = (uint32_t)((want_blocks + nt - 1) / nt);
  uint32_t spc = (uint32_t)((n_slots + chunks - 1) / chunks);
  if (spc < 8) spc = 8;
  if (spc > 512) spc = 512;  // keeps a block around a millisecond: short tails per launch
  chunks = (uint32_t)((n_slots + spc - 1) / spc);

  ScanArgs a;
  memset(&a, 0, sizeof a);
  a.planes = c->d_planes.as<uint32_t>();
  a.text = c->d_text;
  a.n = c->n;
  a.poff = c->d_poff.as<uint32_t>();
  a.plen = c->d_plen.as<uint8_t>();
  a.count = c->d_counters + 0;
  a.seed_count = c->d_counters + 1;
  a.work_count = c->d_counters + 5;
  a.lo = lo;
  a.hi = hi;
  a.tile0 = tile0;
  a.n_planes = (uint32_t)c->n_sym + 1;
  a.n_sym = (uint32_t)c->n_sym;
  a.n_slots = (uint32_t)n_slots;
  a.slots_per_chunk = spc;
  a.n_chunks = chunks;
  a.eos = c->opt.eos;
  a.cls = c->d_cls.as<uint8_t>();
  a.umask = c->d_umask.as<uint64_t>();
  a.n_cls = (uint32_t)c->n_cls;
  a.piece_end = c->d_piece.as<uint64_t>();
  a.narrow = (c->max_len <= 32 && c->n_cls <= NARROW_CLS) ? 1u : 0u;

  const size_t smem = (size_t)(c->n_sym + 1) * TILE_WT * 4;
  IndexArgs ia;
  memset(&ia, 0, sizeof ia);
  void (*ikern)(const IndexArgs) = edit ? pair_scan_kernel<true> : pair_scan_kernel<false>;
  if (c->use_index && !edit) {
    const bool smem_1b = c->idx_proto.log_words2 != 0 && !c->idx_proto.b2_global;
    if (PairKernel sk = pair_sub_kernel((int)c->idx_proto.L, (int)c->idx_proto.nd, smem_1b)) ikern = sk;
  }
  void (*ivkern)(const IndexArgs) =
      K <= 1 ? index_verify_kernel<1> : (K == 2 ? index_verify_kernel<2> : index_verify_kernel<3>);
  if (c->use_index && edit) {
    // the kernel specialised for (piece length, k) and the verifier that reads its seeds;
    // PM_EDIT_GENERIC=1 keeps the run-time kernel pair (for A/B measurements)
    PairKernel ek = c->idx_edit_special ? pair_edit_kernel((int)c->idx_proto.L, K) : nullptr;
    if (c->n >= (1ull << (SEED_POS_BITS - 1)))
      return fail(PM_E_UNSUPPORTED, "shard longer than 2^38 characters for the edit scan of the pair index");
    if (ek) {
      ikern = ek;
      ivkern = K <= 1 ? index_verify_lv_kernel<1> : (K == 2 ? index_verify_lv_kernel<2> : index_verify_lv_kernel<3>);
    }
  }
  size_t ismem = 0;
  if (c->use_index) {
    ia = c->idx_proto;
    ia.pk = c->d_pk.as<uint2>();
    ia.n = c->n;
    ia.n16 = (uint64_t)c->n_tiles * (TILE_POS / 16);
    ia.bitmap = c->d_bitmap.as<uint32_t>();
    ia.bitmap2 = ia.log_words2 ? c->d_bitmap2.as<uint32_t>() : nullptr;
    ia.bucket = c->d_bucket.as<uint32_t>();
    ia.entries = c->d_entries.as<uint2>();
    ia.pacc = c->d_pacc.as<ulonglong4>();
    ia.plen = c->d_plen.as<uint8_t>();
    ia.umask4 = c->d_umask4.as<uint32_t>();
    ia.count = c->d_counters + 0;
    ia.seed_count = c->d_counters + 1;
    ia.work_count = c->d_counters + 5;
    ia.lo = lo;
    ia.hi = hi;
    ia.k = (uint32_t)K;
    ismem = ((size_t)4 << ia.log_words) + (size_t)PAIR_WARPS * 2 * PAIR_QCAP * 8 +
            ((ia.log_words2 && !ia.b2_global) ? ((size_t)4 << ia.log_words2) : 0) +
            (size_t)PAIR_WARPS * 32 * sizeof(uint16_t);   // staging rows of the edit kernels' stage A
    PM_CUDA(cudaFuncSetAttribute(ikern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ismem));
  }
  void (*kern)(const ScanArgs) = nullptr;
  void (*vkern)(const ScanArgs) = nullptr;
  if (edit) {
    switch (K) {
      case 1: kern = edit_seed_kernel<1>; vkern = edit_verify_kernel<1>; break;
      case 2: kern = edit_seed_kernel<2>; vkern = edit_verify_kernel<2>; break;
      case 3: kern = edit_seed_kernel<3>; vkern = edit_verify_kernel<3>; break;
      case 4: kern = edit_seed_kernel<4>; vkern = edit_verify_kernel<4>; break;
      default: kern = edit_seed_kernel<5>; vkern = edit_verify_kernel<5>; break;
    }
  } else {
    switch (K) {
      case 0: kern = hamming_scan_kernel<0>; break;
      case 1: kern = hamming_scan_kernel<1>; break;
      case 2: kern = hamming_scan_kernel<2>; break;
      case 3: kern = hamming_scan_kernel<3>; break;
      case 4: kern = hamming_scan_kernel<4>; break;
      default: kern = hamming_scan_kernel<5>; break;
    }
  }
  PM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  // candidate buffer: one entry per 64 characters of the scanned range (a genome whose repeats hold
  // primer sites gives ~1 hit per kbp), at least 2^20; more than that costs one repeated scan
  uint64_t cap = c->d_keys.cap / 8;
  const uint64_t want_cap = std::max<uint64_t>(1u << 20, (uint64_t)(q_hi - q_lo) / 64);
  if (cap < want_cap) {
    int rc = c->d_keys.ensure((size_t)want_cap * 8);
    if (rc) return rc;
    cap = c->d_keys.cap / 8;
  }
  uint64_t count = 0;
  // edit mode: the seed list of one batch of tiles (dense kernel -> verify kernel)
  uint32_t batch_tiles = nt;
  if (edit) {
    // 2^26 entries unless PM_SEED_LIST_ENTRIES says otherwise (a smaller list on a small GPU share;
    // a batch that does not fit is cut into smaller batches, then the list grows)
    size_t seed_entries = (size_t)64u << 20;
    if (const char *e = getenv("PM_SEED_LIST_ENTRIES")) seed_entries = (size_t)std::max(1024ll, atoll(e));
    int rc = c->d_seeds.ensure(seed_entries * 8);
    if (rc) return rc;
    const double per_tile = std::max(1.0, (c->use_index ? c->idx_seed_rate : c->seed_rate) * (double)TILE_POS);
    const double fit = (double)(c->d_seeds.cap / 8) / (2.0 * per_tile);
    batch_tiles = (uint32_t)std::max(1.0, std::min((double)nt, fit));
  }
  // first-piece positions the index path has to visit: a window starts at most 31 characters earlier
  const int64_t ip_lo = q_lo, ip_hi = std::min<int64_t>((int64_t)c->n, q_hi + 32);
  uint64_t idx_lookups = 0;
  const int kAttempts = 10;
  for (int attempt = 0; attempt < kAttempts; attempt++) {
    idx_lookups = 0;
    ia.out = c->d_keys.as<unsigned long long>();
    ia.cap = cap;
    ia.seeds = c->d_seeds.as<unsigned long long>();
    ia.seed_cap = c->d_seeds.cap / 8;
    ia.seed_n = c->d_counters + 6;
    ia.seed_overflow = c->d_counters + 7;
    a.out = c->d_keys.as<unsigned long long>();
    a.cap = cap;
    a.seeds = c->d_seeds.as<unsigned long long>();
    a.seed_cap = c->d_seeds.cap / 8;
    a.seed_n = c->d_counters + 6;
    a.seed_overflow = c->d_counters + 7;
    PM_CUDA(cudaMemsetAsync(c->d_counters, 0, 2 * sizeof(unsigned long long), c->stream));
    PM_CUDA(cudaMemsetAsync(c->d_counters + 5, 0, 3 * sizeof(unsigned long long), c->stream));
    cudaStream_t ls = c->stream;
    PM_CUDA(cudaEventRecord(c->ev[0], c->stream));
    // grid.x is limited to 2^31-1, grid.y to 65535
    if ((uint64_t)nt * chunks > 0x7fffffffull) return fail(PM_E_UNSUPPORTED, "scan grid too large");
    {
      uint32_t prev_end = 0;
      for (size_t ci = 0; ci < c->chunks.size(); ci++) {
        const pm_ctx::Chunk &ch = c->chunks[ci];
        const uint32_t t0 = std::max(prev_end, tile0), t1 = std::min(ch.tile_end, tile1);
        prev_end = std::max(prev_end, ch.tile_end);
        if (t1 <= t0) continue;
        if (ch.ev) PM_CUDA(cudaStreamWaitEvent(ls, ch.ev, 0));
        if (c->use_index) {
          // a verification window reaches up to 63 characters past the first piece: positions lag
          // the packed tiles by 64 so that every window lies in words that are already packed.
          // The lag is made up by the next chunk's launch; when the range ends inside this chunk
          // (no later chunk is visited) the positions run to the end of the range instead, after
          // the chunks holding the words behind it have arrived.
          const int64_t lag = 64;
          int64_t pb = std::max<int64_t>(ip_lo, (int64_t)t0 * TILE_POS - (t0 ? lag : 0));
          int64_t pe;
          if (ch.tile_end >= c->n_tiles) {
            pe = std::min<int64_t>(ip_hi, (int64_t)c->n);
          } else if (ch.tile_end < tile1) {
            pe = std::min<int64_t>(ip_hi, (int64_t)t1 * TILE_POS - lag);
          } else {
            pe = ip_hi;
            for (size_t cj = ci + 1; cj < c->chunks.size(); cj++) {
              if ((int64_t)c->chunks[cj - 1].tile_end * TILE_POS >= ip_hi + lag) break;
              if (c->chunks[cj].ev) PM_CUDA(cudaStreamWaitEvent(ls, c->chunks[cj].ev, 0));
            }
          }
          const int64_t batch_pos = (int64_t)batch_tiles * TILE_POS;
          for (; pb < pe; pb += batch_pos) {
            ia.p_begin = pb;
            ia.p_end = std::min<int64_t>(pe, pb + batch_pos);
            const uint64_t span = (uint64_t)(ia.p_end - ia.p_begin);
            if (edit) PM_CUDA(cudaMemsetAsync(c->d_counters + 6, 0, sizeof(unsigned long long), ls));
            const uint64_t warp_chunks = (span + 16 + 511) / 512;  // 512 positions per warp pass
            const unsigned iblocks = (unsigned)std::max<uint64_t>(
                1, std::min<uint64_t>((uint64_t)n_sm, (warp_chunks + PAIR_WARPS - 1) / PAIR_WARPS));
            ikern<<<iblocks, PAIR_THREADS, ismem, ls>>>(ia);
            idx_lookups += span * ia.nd;
            PM_CUDA(cudaGetLastError());
            c->stats.kernel_launches++;
            if (edit) {
              ivkern<<<n_sm * 8, 128, 0, ls>>>(ia);
              PM_CUDA(cudaGetLastError());
              c->stats.kernel_launches++;
            }
          }
          continue;
        }
        for (uint32_t b0 = t0; b0 < t1; b0 += batch_tiles) {
          const uint32_t b1 = std::min(t1, b0 + batch_tiles);
          a.tile0 = b0;
          if (edit) PM_CUDA(cudaMemsetAsync(c->d_counters + 6, 0, sizeof(unsigned long long), ls));
          if (edit && b0 == 0) {  // nominal window starts -1 .. -K (see edit_start_seeds_kernel)
            edit_start_seeds_kernel<<<nblocks((uint64_t)n_slots * K, 256), 256, 0, ls>>>(a, K);
            c->stats.kernel_launches++;
          }
          kern<<<(b1 - b0) * chunks, SCAN_THREADS, smem, ls>>>(a);
          PM_CUDA(cudaGetLastError());
          c->stats.kernel_launches++;
          if (edit) {
            vkern<<<n_sm * 8, VERIFY_THREADS, 0, ls>>>(a);
            PM_CUDA(cudaGetLastError());
            c->stats.kernel_launches++;
          }
        }
      }
      // the alignment kernels read the raw text: everything must have arrived
      if (!c->chunks.empty() && c->chunks.back().ev)
        PM_CUDA(cudaStreamWaitEvent(ls, c->chunks.back().ev, 0));
    }
    const bool amb_edit = c->use_index && c->idx_has_amb && edit;
    if (c->use_index && c->idx_has_amb && !edit) {
      // ambiguous text characters (their number is on the device once every chunk is packed):
      // every window around one is decided on the raw text
      amb_sub_kernel<<<(unsigned)n_sm * 16, 256, 0, c->stream>>>(a, c->d_amb.as<unsigned long long>(),
                                                                 c->d_counters + CTR_AMB, kAmbCap, (uint32_t)K);
      PM_CUDA(cudaGetLastError());
      c->stats.kernel_launches++;
    } else if (amb_edit) {
      // edits: the seeds of the ambiguous characters go through the seed list in batches sized
      // on the host (rare: IUPAC text characters under -w together with -k)
      PM_CUDA(cudaMemcpyAsync(c->h_counters + CTR_AMB, c->d_counters + CTR_AMB, sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, c->stream));
      PM_CUDA(cudaStreamSynchronize(c->stream));
      const uint64_t n_amb = c->h_counters[CTR_AMB];
      c->stats.n_ambiguous = n_amb;
      if (n_amb > kAmbCap) return kFallbackToPlanes;
      // seeds per ambiguous character and slot: (m + 2K) / (2K + 1) + 1 alignment-end windows
      const uint64_t per_amb = (uint64_t)n_slots * ((uint64_t)(c->max_len + 2 * K) / (2 * K + 1) + 1);
      const uint64_t batch = std::max<uint64_t>(1, a.seed_cap / std::max<uint64_t>(1, per_amb));
      for (uint64_t a0 = 0; a0 < n_amb; a0 += batch) {
        const uint64_t a1 = std::min(n_amb, a0 + batch);
        PM_CUDA(cudaMemsetAsync(c->d_counters + CTR_SEED_N, 0, sizeof(unsigned long long), c->stream));
        amb_seed_kernel<<<(unsigned)std::min<uint64_t>(a1 - a0, (uint64_t)n_sm * 16), 256, 0, c->stream>>>(
            a, c->d_amb.as<unsigned long long>(), a0, a1, K);
        vkern<<<n_sm * 8, VERIFY_THREADS, 0, c->stream>>>(a);
        PM_CUDA(cudaGetLastError());
        c->stats.kernel_launches += 2;
      }
    }
    PM_CUDA(cudaEventRecord(c->ev[1], c->stream));
    c->stats.scan_items = (uint64_t)nt * chunks;
    c->stats.scan_path = c->use_index ? PM_SCAN_INDEX : PM_SCAN_PLANES;
    c->stats.index_q = (uint32_t)c->idx_q;
    c->stats.index_nd = c->use_index ? c->idx_proto.nd : 0;
    c->stats.index_pieces = c->use_index ? (uint32_t)(K + 2) : 0;
    c->stats.index_lookups = idx_lookups;
    c->stats.index_entries = c->use_index ? c->idx_entries : 0;
    PM_CUDA(cudaMemcpyAsync(c->h_counters, c->d_counters, 10 * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, c->stream));
    PM_CUDA(cudaStreamSynchronize(c->stream));
    c->stats.host_syncs++;
    {
      float ms = 0;
      cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
      c->stats.scan_ms = ms;
    }
    count = c->h_counters[CTR_COUNT];
    c->stats.n_seeds = c->h_counters[CTR_SEEDS];
    c->stats.scan_warp_steps = c->h_counters[CTR_WORK];
    if (c->use_index && c->idx_has_amb) {
      c->stats.n_ambiguous = c->h_counters[CTR_AMB];
      if (c->h_counters[CTR_AMB] > kAmbCap) return kFallbackToPlanes;
    }
    if (edit && c->h_counters[CTR_SEED_OVF]) {
      // a batch produced more seeds than the list holds: smaller batches, then a larger list
      if (batch_tiles > 1) batch_tiles = std::max(1u, batch_tiles / 8);
      else if (int rc = c->d_seeds.ensure(c->d_seeds.cap * 4)) return rc;
      if (attempt == kAttempts - 1) return fail(PM_E_NOMEM, "seed list overflow");
      continue;
    }
    if (count <= cap) break;
    // candidate buffer too small (it starts at one entry per 64 characters of text): the counter
    // kept counting, so one more pass with a buffer of that size is enough
    int rc = c->d_keys.ensure((size_t)count * 8 + 1024);
    if (rc) return rc;
    cap = c->d_keys.cap / 8;
    if (attempt == kAttempts - 1) return fail(PM_E_NOMEM, "candidate buffer overflow");
  }
  if (count == 0) return PM_OK;
  if (c->unordered_ok && K == 0 && c->stats.n_ambiguous == 0) {
    // the exact scan of an exact_halves / exact_bases inner context: every occurrence is found
    // exactly once (one piece pair per pattern) and its consumer orders the records itself
    *n_out = count;
    return PM_OK;
  }

  // sort by (slot, pos); edit mode emits duplicates -> unique
  int rc = c->d_keys2.ensure((size_t)count * 8);
  if (rc) return rc;
  const int end_bit = 40 + slot_bits(n_slots);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, c->d_keys.as<unsigned long long>(),
                                 c->d_keys2.as<unsigned long long>(), count, 0, end_bit, c->stream);
  size_t tmp2 = 0;
  cub::DeviceSelect::Unique(nullptr, tmp2, c->d_keys2.as<unsigned long long>(),
                            c->d_keys.as<unsigned long long>(), c->d_counters + CTR_UNIQ, count, c->stream);
  tmp_bytes = std::max(tmp_bytes, tmp2);
  rc = c->d_tmp.ensure(tmp_bytes);
  if (rc) return rc;
  size_t tb = c->d_tmp.cap;
  PM_CUDA(cub::DeviceRadixSort::SortKeys(c->d_tmp.p, tb, c->d_keys.as<unsigned long long>(),
                                         c->d_keys2.as<unsigned long long>(), count, 0, end_bit,
                                         c->stream));
  c->stats.kernel_launches += 4;
  const bool has_dups = edit || c->use_index;
  if (has_dups) {
    tb = c->d_tmp.cap;
    PM_CUDA(cub::DeviceSelect::Unique(c->d_tmp.p, tb, c->d_keys2.as<unsigned long long>(),
                                      c->d_keys.as<unsigned long long>(), c->d_counters + CTR_UNIQ, count,
                                      c->stream));
    c->stats.kernel_launches += 2;
    if (exact_count) {
      PM_CUDA(cudaMemcpyAsync(c->h_counters + CTR_UNIQ, c->d_counters + CTR_UNIQ, sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, c->stream));
      PM_CUDA(cudaStreamSynchronize(c->stream));
      c->stats.host_syncs++;
      count = c->h_counters[CTR_UNIQ];
    }
  } else {
    std::swap(c->d_keys, c->d_keys2);
    store_u64_kernel<<<1, 1, 0, c->stream>>>(c->d_counters + CTR_UNIQ, count);
    c->stats.kernel_launches++;
  }
  *n_out = count;
  return PM_OK;
}

// sort the first d_counters[CTR_NHITS] records of c->d_hits by their key (c->d_sortk; the keys
// behind them are padding, ~0) and write them in order to `out`: c->d_hits2, or page-locked host
// memory.  `bound` = records + padding, the count the host knows.
int sort_hits(pm_ctx *c, uint64_t bound, pm_hit *out) {
  int rc;
  if ((rc = c->d_sortk2.ensure(bound * 8))) return rc;
  if ((rc = c->d_idx.ensure(bound * 4))) return rc;
  if ((rc = c->d_idx2.ensure(bound * 4))) return rc;
  iota_kernel<<<nblocks(bound, 256), 256, 0, c->stream>>>(c->d_idx.as<uint32_t>(), bound);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, c->d_sortk.as<unsigned long long>(),
                                  c->d_sortk2.as<unsigned long long>(), c->d_idx.as<uint32_t>(),
                                  c->d_idx2.as<uint32_t>(), bound, 0, 64, c->stream);
  if ((rc = c->d_tmp.ensure(tb))) return rc;
  tb = c->d_tmp.cap;
  PM_CUDA(cub::DeviceRadixSort::SortPairs(c->d_tmp.p, tb, c->d_sortk.as<unsigned long long>(),
                                          c->d_sortk2.as<unsigned long long>(),
                                          c->d_idx.as<uint32_t>(), c->d_idx2.as<uint32_t>(), bound, 0,
                                          64, c->stream));
  const unsigned gb = (unsigned)std::min<uint64_t>((bound * HIT_WORDS + 255) / 256, (uint64_t)c->n_sm * 16);
  gather_hits_kernel<<<gb, 256, 0, c->stream>>>(c->d_hits.as<pm_hit>(), c->d_idx2.as<uint32_t>(),
                                                c->d_counters + CTR_NHITS, out);
  PM_CUDA(cudaGetLastError());
  c->stats.kernel_launches += 5;
  return PM_OK;
}

int pack_text(pm_ctx *c);

// run_scan, falling back to the bit-plane kernels when the text holds more ambiguous
// characters than the index path is willing to re-examine one by one
int run_scan_any(pm_ctx *c, int64_t lo, int64_t hi, uint64_t *n_out, bool exact_count) {
  int rc = run_scan(c, lo, hi, n_out, exact_count);
  if (rc != kFallbackToPlanes) return rc;
  if (c->scan_path == PM_SCAN_INDEX)
    return fail(PM_E_UNSUPPORTED, "too many ambiguous text characters for the pair-seed index scan");
  c->use_index = false;  // until the next pm_set_text_*
  const uint64_t n_amb = c->stats.n_ambiguous;
  if ((rc = pack_text(c))) return rc;
  rc = run_scan(c, lo, hi, n_out, exact_count);
  c->stats.n_ambiguous = n_amb;
  return rc;
}

int find_patterns_guarded(pm_ctx *c, int64_t begin, int64_t end, uint64_t *nhits, int64_t right_guard, bool to_host);

// The hits whose chain head lies in [begin, end).  Chains are followed past `end` into a right
// guard; when one is still alive at the end of the guard the scan is repeated with a guard
// eight times as long (up to the end of the loaded text).  to_host: the sorted records are
// written to c->h_out (page-locked), otherwise to c->d_hits2.
int find_patterns_impl(pm_ctx *c, int64_t begin, int64_t end, uint64_t *nhits, bool to_host) {
  int64_t guard = kRightGuard;
  for (;;) {
    const int rc = find_patterns_guarded(c, begin, end, nhits, guard, to_host);
    if (rc != kGuardTooShort) return rc;
    guard *= 8;
  }
}

// Host synchronisations of one call: the scan's counters (run_scan) and the final read-back of
// every other count together with the hit records.  Everything in between is sized by `bound`,
// the number of raw candidates, and reads its real counts from d_counters.
int find_patterns_guarded(pm_ctx *c, int64_t begin, int64_t end, uint64_t *nhits, int64_t right_guard, bool to_host) {
  *nhits = 0;
  if (!c->inited) return fail(PM_E_ARG, "pm_find_patterns before pm_init");
  if (!c->d_text && c->n) return fail(PM_E_ARG, "no text");
  memset(&c->stats, 0, offsetof(pm_stats, pack_ms));
  c->stats.text_len = c->n;
  c->stats.n_patterns = c->pats.size();
  PM_CUDA(cudaSetDevice(c->opt.device));
  PM_CUDA(cudaEventRecord(c->ev[2], c->stream));
  const int K = c->opt.k;
  // local candidate-position window
  int64_t own_lo = begin - c->offset, own_hi = end - c->offset;
  if (begin < c->offset) own_lo = 0;
  if (own_lo < 0) own_lo = 0;
  if (end == std::numeric_limits<int64_t>::max() || own_hi > (int64_t)c->n + 1) own_hi = (int64_t)c->n + 1;
  if (own_hi <= own_lo) return PM_OK;
  int64_t lo = own_lo, hi = own_hi;
  bool guarded = false;
  if (K > 0) {
    lo = own_lo - (2 * K + 1);
    if (lo < 0) lo = 0;
    if (own_hi < (int64_t)c->n + 1) {
      hi = own_hi + right_guard;
      guarded = true;
      if (hi >= (int64_t)c->n + 1) { hi = (int64_t)c->n + 1; guarded = false; }
    }
  }
  uint64_t bound = 0;   // candidates: an upper bound on every later count
  int rc;
  if (c->inner) {
    // exact_halves: every exact occurrence of a pattern half up to the end of the range (the
    // `lasthit` rule looks back, so the scan starts where the loaded text starts)
    guarded = false;
    memset(&c->inner->stats, 0, offsetof(pm_stats, pack_ms));   // the inner context's counters of THIS call
    rc = run_scan_any(c->inner, 0, own_hi, &bound, true);
    const pm_stats &is = c->inner->stats;
    c->stats.scan_ms = is.scan_ms;
    c->stats.scan_path = is.scan_path;
    c->stats.scan_items = is.scan_items;
    c->stats.scan_warp_steps = is.scan_warp_steps;
    c->stats.index_q = is.index_q;
    c->stats.index_nd = is.index_nd;
    c->stats.index_pieces = is.index_pieces;
    c->stats.index_lookups = is.index_lookups;
    c->stats.index_entries = is.index_entries;
    c->stats.n_ambiguous = is.n_ambiguous;
    c->stats.kernel_launches += is.kernel_launches;
    c->stats.host_syncs += is.host_syncs;
  } else {
    rc = run_scan_any(c, lo, hi, &bound, false);
  }
  if (rc) return rc;
  c->stats.n_candidates = bound;
  unsigned long long *const dc = c->d_counters;
  bool have_tail = false;
  if (bound > 0) {
    have_tail = true;
    const AlignEnv env = make_env(c);
    const unsigned long long *keys = c->d_keys.as<unsigned long long>();
    // where the sorted records go: page-locked host memory the kernels write directly (no count
    // needed on the host), as long as the bound is moderate
    const bool direct = to_host && bound * sizeof(pm_hit) <= kDirectOutBytes;
    if (direct && (rc = ensure_host_out(c, bound * sizeof(pm_hit)))) return rc;
    if (!direct && (rc = c->d_hits2.ensure(bound * sizeof(pm_hit)))) return rc;
    pm_hit *const out = direct ? reinterpret_cast<pm_hit *>(c->h_out) : c->d_hits2.as<pm_hit>();
    if ((rc = c->d_hits.ensure(bound * sizeof(pm_hit)))) return rc;
    if ((rc = c->d_sortk.ensure(bound * 8))) return rc;
    // counters of the tail: bogus, raw hits / cut flag / halves / chains, pool, hits
    PM_CUDA(cudaMemsetAsync(dc + CTR_BOGUS, 0, 2 * sizeof(unsigned long long), c->stream));
    PM_CUDA(cudaMemsetAsync(dc + CTR_CUT, 0, 2 * sizeof(unsigned long long), c->stream));
    PM_CUDA(cudaMemsetAsync(dc + CTR_CHAINS, 0, 3 * sizeof(unsigned long long), c->stream));
    const unsigned long long *n_final = dc + CTR_RAW;   // hits that get a final alignment
    if (c->inner) {
      // extend every half occurrence over the other half, order them as exact_halves.cc:113-116
      // does, apply its lasthit rule per pattern
      if ((rc = c->d_half_rec.ensure(bound * sizeof(HalfRec)))) return rc;
      if ((rc = c->d_half_key.ensure(bound * 8))) return rc;
      if ((rc = c->d_half_key2.ensure(bound * 8))) return rc;
      if ((rc = c->d_half_ord.ensure(bound * 4))) return rc;
      if ((rc = c->d_half_ord2.ensure(bound * 4))) return rc;
      if ((rc = c->d_raw.ensure(bound * sizeof(RawHit)))) return rc;
      // the occurrence count is exact here (run_scan read it back); kept in the inner counters
      const unsigned long long *n_occ = c->inner->d_counters + CTR_UNIQ;
      store_u64_kernel<<<1, 1, 0, c->stream>>>(c->inner->d_counters + CTR_UNIQ, bound);
      if (!c->inner_bases)   // padding of the sort below
        PM_CUDA(cudaMemsetAsync(c->d_half_key.p, 0xff, bound * 8, c->stream));
      halves_extend_kernel<<<nblocks(bound, 128), 128, 0, c->stream>>>(
          env, c->inner->d_keys.as<unsigned long long>(), n_occ, bound, c->inner->d_slotid.as<uint32_t>(),
          c->inner_bases ? 1 : 0, c->d_half_rec.as<HalfRec>(), c->d_half_key.as<unsigned long long>(),
          dc + CTR_HALF_OK, own_lo, own_hi, c->d_raw.as<RawHit>(), dc + CTR_RAW, bound);
      PM_CUDA(cudaGetLastError());
      c->stats.kernel_launches += 2;
      if (!c->inner_bases) {
        // (exact_bases: every extension is a hit)
        iota_kernel<<<nblocks(bound, 256), 256, 0, c->stream>>>(c->d_half_ord.as<uint32_t>(), bound);
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, c->d_half_key.as<unsigned long long>(),
                                        c->d_half_key2.as<unsigned long long>(), c->d_half_ord.as<uint32_t>(),
                                        c->d_half_ord2.as<uint32_t>(), bound, 0, 64, c->stream);
        if ((rc = c->d_tmp.ensure(tb))) return rc;
        tb = c->d_tmp.cap;
        PM_CUDA(cub::DeviceRadixSort::SortPairs(c->d_tmp.p, tb, c->d_half_key.as<unsigned long long>(),
                                                c->d_half_key2.as<unsigned long long>(),
                                                c->d_half_ord.as<uint32_t>(), c->d_half_ord2.as<uint32_t>(), bound,
                                                0, 64, c->stream));
        const uint32_t n_slots = (uint32_t)c->pats.size();
        halves_report_kernel<<<nblocks(n_slots, 128), 128, 0, c->stream>>>(
            c->d_half_key2.as<unsigned long long>(), c->d_half_ord2.as<uint32_t>(), c->d_half_rec.as<HalfRec>(),
            dc + CTR_HALF_OK, n_slots, K, c->opt.indels, own_lo, own_hi, c->d_raw.as<RawHit>(), dc + CTR_RAW, bound);
        PM_CUDA(cudaGetLastError());
        c->stats.kernel_launches += 5;
      }
    } else if (K == 0) {
      exact_hits_kernel<<<nblocks(bound, 128), 128, 0, c->stream>>>(
          env, keys, dc + CTR_UNIQ, bound, c->d_slotid.as<uint32_t>(), c->offset, c->d_hits.as<pm_hit>(),
          c->d_sortk.as<unsigned long long>(), dc + CTR_BOGUS, dc + CTR_NHITS, c->align_rules == PM_ALIGN_PCR_MATCH);
      PM_CUDA(cudaGetLastError());
      c->stats.kernel_launches++;
      n_final = nullptr;
    } else {
      // chains of candidates (filter_bitvec.cc:157-168) and their yes/no alignment
      if ((rc = c->d_head.ensure(bound * 4))) return rc;
      if ((rc = c->d_incl.ensure(bound * 4))) return rc;
      if ((rc = c->d_chains.ensure(bound * sizeof(Chain)))) return rc;
      if ((rc = c->d_raw.ensure(bound * sizeof(RawHit)))) return rc;
      chain_heads2_kernel<<<nblocks(bound, 256), 256, 0, c->stream>>>(keys, dc + CTR_UNIQ, bound, 2 * K + 1,
                                                                     c->d_head.as<uint32_t>());
      size_t tb = 0;
      cub::DeviceScan::InclusiveSum(nullptr, tb, c->d_head.as<uint32_t>(), c->d_incl.as<uint32_t>(),
                                    bound, c->stream);
      if ((rc = c->d_tmp.ensure(tb))) return rc;
      tb = c->d_tmp.cap;
      PM_CUDA(cub::DeviceScan::InclusiveSum(c->d_tmp.p, tb, c->d_head.as<uint32_t>(),
                                            c->d_incl.as<uint32_t>(), bound, c->stream));
      chain_fill2_kernel<<<nblocks(bound, 256), 256, 0, c->stream>>>(
          keys, dc + CTR_UNIQ, c->d_head.as<uint32_t>(), c->d_incl.as<uint32_t>(), c->d_chains.as<Chain>(),
          dc + CTR_CHAINS);
      if (!c->d_pool.p && (rc = c->d_pool.ensure(kPoolBytes))) return rc;
      TailCounts tc;
      tc.n_cands = dc + CTR_UNIQ;
      tc.n_chains = dc + CTR_CHAINS;
      tc.n_raw = dc + CTR_RAW;
      tc.n_bogus = dc + CTR_BOGUS;
      tc.cut = dc + CTR_CUT;
      tc.pool_used = dc + CTR_POOL;
      chain_align2_kernel<<<nblocks(bound, 64), 64, 0, c->stream>>>(
          env, c->d_chains.as<Chain>(), tc, bound, own_lo, own_hi, guarded ? hi : 0, c->d_pool.as<uint8_t>(),
          (unsigned long long)c->d_pool.cap, c->d_raw.as<RawHit>());
      PM_CUDA(cudaGetLastError());
      c->stats.kernel_launches += 5;
    }
    if (n_final) {
      final_align_kernel<<<nblocks(bound, 64), 64, 0, c->stream>>>(
          env, c->d_raw.as<RawHit>(), n_final, bound, c->d_slotid.as<uint32_t>(), c->offset,
          c->d_hits.as<pm_hit>(), c->d_sortk.as<unsigned long long>(), dc + CTR_BOGUS, dc + CTR_NHITS,
          c->inner ? 1 : 0);
      PM_CUDA(cudaGetLastError());
      c->stats.kernel_launches++;
    }
    if ((rc = sort_hits(c, bound, out))) return rc;
    c->out_direct = direct;
  }
  // the one read-back: every count of the tail (the records are on their way to `out` already)
  PM_CUDA(cudaMemcpyAsync(c->h_counters, c->d_counters, CTR_N * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, c->stream));
  PM_CUDA(cudaEventRecord(c->ev[3], c->stream));
  PM_CUDA(cudaEventSynchronize(c->ev[3]));
  c->stats.host_syncs++;
  c->last_uniq = (bound > 0 && !c->inner) ? c->h_counters[CTR_UNIQ] : 0;
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
    c->stats.total_ms = ms;
    if (c->pack_timed && cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]) == cudaSuccess)
      c->stats.pack_ms = ms;
    else
      cudaGetLastError();
  }
  if (have_tail) {
    const unsigned long long *h = c->h_counters;
    if (!c->inner && K > 0) {
      if (h[CTR_POOL] > c->d_pool.cap) {
        // chains longer than the per-thread band (repeats) asked for more pool than there is:
        // grow it and run the call again (the scan included: rare, and the pool stays)
        if ((rc = c->d_pool.ensure((size_t)h[CTR_POOL] + (1u << 20)))) return rc;
        return find_patterns_guarded(c, begin, end, nhits, right_guard, to_host);
      }
      if (guarded && h[CTR_CUT]) return kGuardTooShort;  // an owned chain reaches the end of the right guard
      c->stats.n_chains = h[CTR_CHAINS];
    }
    if (!c->inner) c->stats.n_candidates = h[CTR_UNIQ];
    if (c->inner) c->stats.n_chains = c->inner_bases ? h[CTR_RAW] : h[CTR_HALF_OK];
    const uint64_t n_out = h[CTR_NHITS];
    c->stats.n_bogus = h[CTR_BOGUS];
    c->stats.n_hits = n_out;
    *nhits = n_out;
    if (to_host && !c->out_direct && n_out) {
      // more candidates than the direct window holds: the records were sorted into device
      // memory; copy exactly the hits
      if ((rc = ensure_host_out(c, n_out * sizeof(pm_hit)))) return rc;
      PM_CUDA(cudaMemcpyAsync(c->h_out, c->d_hits2.p, n_out * sizeof(pm_hit), cudaMemcpyDeviceToHost, c->stream));
      PM_CUDA(cudaStreamSynchronize(c->stream));
      c->stats.host_syncs++;
    }
  }
  if (c->stats.n_bogus)
    return fail(PM_E_BOGUS, "Bogus hit returned to primer_match main()");
  return PM_OK;
}

cudaEvent_t next_event(pm_ctx *c) {
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    c->ev_pool.push_back(e);
  }
  return c->ev_pool[c->ev_used++];
}

int pack_tiles(pm_ctx *c, uint32_t t0, uint32_t t1, cudaStream_t st) {
  if (t1 <= t0) return PM_OK;
  if (c->use_index) {
    const uint64_t w0 = (uint64_t)t0 * (TILE_POS / 16), w1 = (uint64_t)t1 * (TILE_POS / 16);
    const uint64_t per_block = 256ull * PACK_WORDS_PER_THREAD;
    pack_linear_kernel<<<(unsigned)((w1 - w0 + per_block - 1) / per_block), 256, 0, st>>>(
        c->d_text, c->n, c->d_lin.as<uint8_t>(), c->opt.eos, c->d_pk.as<uint2>(), w0, w1,
        c->d_amb.as<unsigned long long>(), c->d_counters + 8, c->d_amb.p ? kAmbCap : 0,
        c->idx_dna_fast ? 1 : 0);
    PM_CUDA(cudaGetLastError());
    return PM_OK;
  }
  pack_planes_kernel<<<t1 - t0, PACK_THREADS, 0, st>>>(c->d_text, c->n, c->d_symmask.as<uint32_t>(),
                                                       c->n_sym + 1, c->opt.eos,
                                                       c->d_planes.as<uint32_t>(), t0);
  PM_CUDA(cudaGetLastError());
  return PM_OK;
}

int prepare_planes(pm_ctx *c) {
  c->n_tiles = (uint32_t)((c->n + TILE_POS - 1) / TILE_POS);
  c->chunks.clear();
  c->ev_used = 0;
  c->pack_timed = false;
  if (c->n_tiles == 0) return PM_OK;
  if (c->use_index) {
    if (c->idx_has_amb)
      if (int rc = c->d_amb.ensure(kAmbCap * sizeof(unsigned long long))) return rc;
    return c->d_pk.ensure((size_t)c->n_tiles * (TILE_POS / 16) * sizeof(uint2));
  }
  return c->d_planes.ensure((size_t)c->n_tiles * (c->n_sym + 1) * TILE_WT * 4);
}

// whole text already on the device: one pack launch on the main stream.  (Packing chunk
// by chunk on a second stream under the scan of the previous chunk was tried -- 28-warp scan
// CTAs leave registers for one pack block per SM -- and gained nothing: DESIGN.md.)
int pack_text(pm_ctx *c) {
  int rc = prepare_planes(c);
  if (rc || c->n_tiles == 0) return rc;
  PM_CUDA(cudaEventRecord(c->ev[4], c->stream));
  PM_CUDA(cudaMemsetAsync(c->d_counters + 8, 0, sizeof(unsigned long long), c->stream));
  if ((rc = pack_tiles(c, 0, c->n_tiles, c->stream))) return rc;
  PM_CUDA(cudaEventRecord(c->ev[5], c->stream));
  c->pack_timed = true;
  c->chunks.push_back(pm_ctx::Chunk{c->n_tiles, nullptr});
  return PM_OK;
}

}  // namespace

// ===========================================================================
extern "C" {

const char *pm_last_error(void) { return g_err.c_str(); }
const char *pm_version(void) { return "pm_b200 0.1 (sm_100a)"; }

int pm_create(pm_ctx **out, const pm_options *opt) {
  if (!out || !opt) return fail(PM_E_ARG, "null argument");
  *out = nullptr;
  if (opt->k < 0 || opt->k > PM_MAX_K) return fail(PM_E_UNSUPPORTED, "k out of range (0..5)");
  int index = opt->index;
  bool index_auto = false;
  if (index == PM_INDEX_AUTO) {
    if (opt->k > 0) { index = PM_INDEX_FILTER_BITVEC; index_auto = true; }  // or exact_halves: pm_init decides
    else index = opt->wildcards ? PM_INDEX_SHIFT_AND : PM_INDEX_KEYWORD_TREE;
  }
  switch (index) {
    case PM_INDEX_KEYWORD_TREE:
    case 1: case 2:  // keyword_tree with other node types: same hit list
      if (opt->k != 0) return fail(PM_E_ARG, "keyword_tree is an exact index (k must be 0)");
      if (opt->wildcards) return fail(PM_E_ARG, "keyword_tree does not honour wildcards");
      index = PM_INDEX_KEYWORD_TREE;
      break;
    case PM_INDEX_SHIFT_AND:
      if (opt->k != 0) return fail(PM_E_ARG, "shift_and is an exact index (k must be 0)");
      break;
    case PM_INDEX_FILTER_BITVEC:
      if (opt->k == 0) return fail(PM_E_ARG, "filter_bitvec needs k > 0");
      break;
    case 7: case 8: case PM_INDEX_EXACT_BASES: case 10:      // over any of the exact indexes: same hit list
      if (opt->k == 0) return fail(PM_E_ARG, "exact_bases needs k > 0");
      index = PM_INDEX_EXACT_BASES;
      break;
    case 11: case 12: case PM_INDEX_EXACT_HALVES: case 14:  // over any of the exact indexes: same hit list
      if (opt->k == 0) return fail(PM_E_ARG, "exact_halves needs k > 0");
      index = PM_INDEX_EXACT_HALVES;
      break;
    default:
      return fail(PM_E_UNSUPPORTED, "index selection not implemented");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PM_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (opt->device < 0 || opt->device >= ndev) return fail(PM_E_ARG, "bad device ordinal");
  PM_CUDA(cudaSetDevice(opt->device));
  pm_ctx *c = new pm_ctx();
  c->opt = *opt;
  c->opt.indels = opt->indels ? 1 : 0;
  c->opt.wildcards = opt->wildcards ? 1 : 0;
  c->opt.text_n = opt->text_n ? 1 : 0;
  c->index = index;
  c->index_auto = index_auto;
  memset(&c->stats, 0, sizeof c->stats);
  if (cudaMalloc((void **)&c->d_counters, CTR_N * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMallocHost((void **)&c->h_counters, CTR_N * sizeof(unsigned long long)) != cudaSuccess) {
    delete c;
    return fail(PM_E_CUDA, "allocation of counters failed");
  }
  cudaMemset(c->d_counters, 0, CTR_N * sizeof(unsigned long long));
  cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, opt->device);
  for (int i = 0; i < 6; i++) cudaEventCreate(&c->ev[i]);
  for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming);
  cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&c->prep_stream, cudaStreamNonBlocking);

  *out = c;
  return PM_OK;
}

void pm_destroy(pm_ctx *c) {
  if (!c) return;
  if (c->inner) pm_destroy(c->inner);
  c->inner = nullptr;
  cudaSetDevice(c->opt.device);
  cudaDeviceSynchronize();
  delete c->reader;
  c->reader = nullptr;
  DevBuf *bufs[] = {&c->d_symmask, &c->d_poff,   &c->d_plen,   &c->d_patchars, &c->d_lconst,
                    &c->d_rconst,  &c->d_compat, &c->d_cls,    &c->d_umask,    &c->d_piece,
                    &c->d_slotid,  &c->d_text_own, &c->d_planes, &c->d_keys,   &c->d_keys2,
                    &c->d_head,    &c->d_incl,   &c->d_chains,
                    &c->d_raw,    &c->d_hits,   &c->d_hits2,    &c->d_sortk,
                    &c->d_sortk2,  &c->d_idx,    &c->d_idx2,   &c->d_tmp,      &c->d_cands, &c->d_seeds, &c->d_pk,
                    &c->d_bucket,  &c->d_entries, &c->d_pacc,  &c->d_umask4,   &c->d_lin, &c->d_bitmap, &c->d_bitmap2, &c->d_amb, &c->d_sqz, &c->d_sqztab,
                    &c->d_fa_raw, &c->d_fa_map, &c->d_fa_scan, &c->d_fa_cnt, &c->d_fa_chdr, &c->d_fa_start,
                    &c->d_fa_nout, &c->d_fa_nhdr, &c->d_fa_ooff, &c->d_fa_hoff, &c->d_fa_key, &c->d_fa_hs,
                    &c->d_fa_he, &c->d_fa_tmp, &c->d_pool, &c->d_half_rec, &c->d_half_key, &c->d_half_key2, &c->d_half_ord, &c->d_half_ord2};
  for (DevBuf *b : bufs) b->release();
  if (c->d_counters) cudaFree(c->d_counters);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->h_out) cudaFreeHost(c->h_out);
  for (int i = 0; i < 2; i++) {
    if (c->stage[i]) cudaFreeHost(c->stage[i]);
    if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
  }
  for (int i = 0; i < 6; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
  delete c;
}

int pm_set_alignment_rules(pm_ctx *c, int rules) {
  if (!c) return fail(PM_E_ARG, "null context");
  if (rules != PM_ALIGN_PRIMER_MATCH && rules != PM_ALIGN_PCR_MATCH) return fail(PM_E_ARG, "bad alignment rules");
  c->align_rules = rules;
  return PM_OK;
}

int pm_set_alphabet_size(pm_ctx *c, int size) {
  if (!c) return fail(PM_E_ARG, "null context");
  if (c->inited) return fail(PM_E_ARG, "pm_set_alphabet_size after pm_init");
  if (size < 1 || size > 256) return fail(PM_E_ARG, "alphabet size must be 1..256");
  c->alphabet_size = size;
  return PM_OK;
}

int pm_set_scan_path(pm_ctx *c, int path) {
  if (!c) return fail(PM_E_ARG, "null context");
  if (c->inited) return fail(PM_E_ARG, "pm_set_scan_path after pm_init");
  if (path != PM_SCAN_AUTO && path != PM_SCAN_PLANES && path != PM_SCAN_INDEX)
    return fail(PM_E_ARG, "bad scan path");
  c->scan_path = path;
  return PM_OK;
}

int pm_set_stream(pm_ctx *c, void *cuda_stream) {
  if (!c) return fail(PM_E_ARG, "null context");
  c->stream = (cudaStream_t)cuda_stream;
  if (c->inner) c->inner->stream = c->stream;
  return PM_OK;
}

int pm_add_pattern(pm_ctx *c, const char *pattern, uint32_t id, int esb, int eeb) {
  if (!c || !pattern) return fail(PM_E_ARG, "null argument");
  if (c->inited) return fail(PM_E_ARG, "pm_add_pattern after pm_init");
  if (id == 0 || id >= (1u << 24)) return fail(PM_E_ARG, "pattern id must be in 1..2^24-1");
  const size_t m = strlen(pattern);
  if (m == 0) return fail(PM_E_ARG, "empty pattern");
  if (m > PM_MAX_PATTERN_LEN) return fail(PM_E_UNSUPPORTED, "pattern longer than PM_MAX_PATTERN_LEN");
  if ((int)m <= c->opt.k)
    return fail(PM_E_ARG, "Number of edits >= Minimum number of inexact bases");  // select.cc:88-91
  for (size_t i = 0; i < m; i++)
    if ((unsigned char)pattern[i] == c->opt.eos)
      return fail(PM_E_ARG, "pattern contains the end-of-sequence character");
  c->pats.push_back(Pattern{std::string(pattern), id, esb, eeb});
  return PM_OK;
}

int pm_init(pm_ctx *c) {
  if (!c) return fail(PM_E_ARG, "null context");
  if (c->inited) return fail(PM_E_ARG, "pm_init called twice");
  PM_CUDA(cudaSetDevice(c->opt.device));
  const size_t ns = c->pats.size();
  const bool wc = c->opt.wildcards, tn = c->opt.text_n;
  // symbols = distinct pattern characters
  int sym_of[256];
  for (int i = 0; i < 256; i++) sym_of[i] = -1;
  std::vector<unsigned char> sym_char;
  c->max_len = 0;
  for (const Pattern &p : c->pats) {
    c->max_len = std::max<int>(c->max_len, (int)p.s.size());
    for (unsigned char ch : p.s)
      if (sym_of[ch] < 0) {
        sym_of[ch] = (int)sym_char.size();
        sym_char.push_back(ch);
      }
  }
  if ((int)sym_char.size() > MAX_SYMBOLS)
    return fail(PM_E_UNSUPPORTED, "more than 30 distinct pattern characters");
  c->n_sym = (int)sym_char.size();
  std::vector<uint32_t> symmask(256, 0);
  for (int b = 0; b < 256; b++) {
    uint32_t msk = 0;
    for (int s = 0; s < c->n_sym; s++)
      if (scan_match(sym_char[s], (unsigned)b, wc, tn)) msk |= 1u << s;
    if (b == c->opt.eos) msk |= 1u << c->n_sym;
    symmask[b] = msk;
  }
  // text byte classes = distinct symbol masks (for the per-seed verifier)
  std::vector<uint8_t> cls(256, 0);
  std::vector<uint32_t> cls_mask;
  for (int b = 0; b < 256; b++) {
    const uint32_t msk = symmask[b] & ((1u << c->n_sym) - 1u);
    size_t j = 0;
    for (; j < cls_mask.size(); j++)
      if (cls_mask[j] == msk) break;
    if (j == cls_mask.size()) cls_mask.push_back(msk);
    cls[b] = (uint8_t)j;
  }
  c->n_cls = (int)cls_mask.size();

  // Slot order: patterns sorted by their first four characters, so that the dense
  // scan computes the counter state of a shared 4-character prefix once per run
  // (PLEN_SHARE).  Slots are internal; hits carry the caller's pattern ids.
  std::vector<size_t> order(ns);
  for (size_t i = 0; i < ns; i++) order[i] = i;
  auto prefix_key = [&](size_t i) {
    const std::string &ps = c->pats[i].s;
    uint32_t key = 0;
    for (int j = 0; j < 4; j++)
      key = key * 64u + (j < (int)ps.size() ? (uint32_t)sym_of[(unsigned char)ps[j]] + 1u : 0u);
    return key;
  };
  std::stable_sort(order.begin(), order.end(),
                   [&](size_t x, size_t y) { return prefix_key(x) < prefix_key(y); });
  {
    std::vector<Pattern> sorted;
    sorted.reserve(ns);
    for (size_t i = 0; i < ns; i++) sorted.push_back(c->pats[order[i]]);
    c->pats.swap(sorted);
  }

  std::vector<uint32_t> poff((ns + 2) * POFF_STRIDE, (uint32_t)c->n_sym * TILE_WT * 4u);
  std::vector<uint8_t> plen(ns + 1, 0), patchars((ns + 1) * PAT_STRIDE, 0);
  std::vector<uint64_t> piece(ns + 1, 0);
  std::vector<int16_t> lconst(ns + 1, 0), rconst(ns + 1, 0);
  std::vector<uint32_t> slotid(ns + 1, 0);
  std::vector<uint64_t> umask((ns + 1) * (size_t)c->n_cls, 0);
  const int K = c->opt.k;
  for (size_t i = 0; i < ns; i++) {
    const Pattern &p = c->pats[i];
    const int m = (int)p.s.size();
    plen[i] = (uint8_t)m;
    slotid[i] = p.id;
    lconst[i] = (int16_t)p.esb;
    rconst[i] = (int16_t)p.eeb;
    for (int j = 0; j < m; j++) {
      const int s = sym_of[(unsigned char)p.s[j]];
      poff[i * POFF_STRIDE + j] = (uint32_t)(s * TILE_WT + j) * 4u;
      patchars[i * PAT_STRIDE + j] = (uint8_t)p.s[j];
      for (int cl = 0; cl < c->n_cls; cl++)
        if ((cls_mask[cl] >> s) & 1u) umask[i * c->n_cls + cl] |= 1ull << j;
    }
    if (i > 0 && memcmp(&poff[i * POFF_STRIDE], &poff[(i - 1) * POFF_STRIDE], 4 * sizeof(uint32_t)) == 0)
      plen[i] |= (uint8_t)PLEN_SHARE;
    // k+1 pieces, as even as possible (pigeonhole partition)
    for (int j = 1; j <= K + 1; j++) {
      const int64_t e0 = (int64_t)m * (j - 1) / (K + 1), e1 = (int64_t)m * j / (K + 1);
      piece[i] |= 1ull << (e1 - 1);
      c->seed_rate += std::pow(0.25, (double)(e1 - e0));
    }
  }
  // iupac_compatible as a 128x128 bit matrix
  std::vector<uint32_t> cbits(128 * 4, 0);
  for (int r = 0; r < 128; r++)
    for (int col = 0; col < 128; col++)
      if (compat().ok[r][col]) cbits[r * 4 + (col >> 5)] |= 1u << (col & 31);

  struct Up { DevBuf *b; const void *src; size_t bytes; };
  const Up ups[] = {{&c->d_symmask, symmask.data(), symmask.size() * 4},
                    {&c->d_poff, poff.data(), poff.size() * 4},
                    {&c->d_plen, plen.data(), plen.size()},
                    {&c->d_patchars, patchars.data(), patchars.size()},
                    {&c->d_lconst, lconst.data(), lconst.size() * 2},
                    {&c->d_rconst, rconst.data(), rconst.size() * 2},
                    {&c->d_compat, cbits.data(), cbits.size() * 4},
                    {&c->d_cls, cls.data(), cls.size()},
                    {&c->d_umask, umask.data(), umask.size() * 8},
                    {&c->d_piece, piece.data(), piece.size() * 8},
                    {&c->d_slotid, slotid.data(), slotid.size() * 4}};
  for (const Up &u : ups) {
    int rc = u.b->ensure(u.bytes);
    if (rc) return rc;
    PM_CUDA(cudaMemcpy(u.b->p, u.src, u.bytes, cudaMemcpyHostToDevice));
  }
  // ---- pick_pattern_index for k == 1 (select.cc:120-126): exact_halves ---------------------
  if (c->index_auto && K == 1 && ns > 0) {
    long min_length = 1 << 30, cumbooldiff = 0, cumdiff = 0;
    for (const Pattern &p : c->pats) {
      const long len = (long)p.s.size(), ex = std::max(p.esb, p.eeb);
      min_length = std::min(min_length, len);
      cumdiff += ex - len / 2;
      cumbooldiff += (ex - len / 2 >= 0) ? 1 : 0;
    }
    // select.cc:120-123: the alphabet size of the database decides how long the shortest primer
    // has to be.  A streamed FASTA database has CharacterProducer::size() == 256 (char_io.h:96),
    // a .sqn / .sqz database the size of its character table (pm_set_alphabet_size).
    const long need = (c->alphabet_size < 10) ? 12 : 8;
    if (min_length >= need && (cumbooldiff <= 0 || cumdiff <= 0)) c->index = PM_INDEX_EXACT_HALVES;
  }
  if (c->index_auto && K > 0 && ns > 0 && c->index == PM_INDEX_FILTER_BITVEC) {
    // select.cc:131-133: exact_bases when every pattern has 6 or more constrained characters
    int min_exact_const = 1 << 30;
    for (const Pattern &p : c->pats) min_exact_const = std::min(min_exact_const, std::max(p.esb, p.eeb));
    if (min_exact_const >= 6) c->index = PM_INDEX_EXACT_BASES;
  }
  if (K > 0)  // select.cc:84-91
    for (const Pattern &p : c->pats)
      if ((int)p.s.size() - std::max(p.esb, p.eeb) <= K && std::max(p.esb, p.eeb) > 0)
        return fail(PM_E_ARG, "Number of edits >= Minimum number of inexact bases");
  if (c->index == PM_INDEX_EXACT_BASES) {
    // the exact index over the constrained prefix (esb >= eeb) or suffix of every pattern;
    // inner pattern id = slot + 1 (exact_bases.cc:126-159)
    pm_options io = c->opt;
    io.k = 0;
    io.indels = 0;
    io.index = c->opt.wildcards ? PM_INDEX_SHIFT_AND : PM_INDEX_KEYWORD_TREE;
    int rc = pm_create(&c->inner, &io);
    if (rc) return rc;
    c->inner_bases = true;
    c->inner->unordered_ok = true;
    c->inner->scan_path = (c->scan_path == PM_SCAN_PLANES) ? PM_SCAN_PLANES : PM_SCAN_AUTO;
    c->inner->stream = c->stream;
    for (size_t i = 0; i < ns; i++) {
      const Pattern &p = c->pats[i];
      const int m = (int)p.s.size();
      const int len = std::max(p.esb, p.eeb);
      if (len <= 0 || len > m) return fail(PM_E_ARG, "exact_bases needs an exact-base constraint on every pattern");
      const std::string part = (p.esb >= p.eeb) ? p.s.substr(0, (size_t)p.esb) : p.s.substr((size_t)(m - p.eeb));
      if ((rc = pm_add_pattern(c->inner, part.c_str(), (uint32_t)(i + 1), 0, 0))) return rc;
    }
    if ((rc = pm_init(c->inner))) return rc;
  }
  if (c->index == PM_INDEX_EXACT_HALVES) {
    // the exact index over the halves: inner pattern ids 2i+1 (left), 2i+2 (right) for slot i
    // (exact_halves.cc:201-224)
    pm_options io = c->opt;
    io.k = 0;
    io.indels = 0;
    io.index = c->opt.wildcards ? PM_INDEX_SHIFT_AND : PM_INDEX_KEYWORD_TREE;
    int rc = pm_create(&c->inner, &io);
    if (rc) return rc;
    c->inner->unordered_ok = true;
    c->inner->scan_path = (c->scan_path == PM_SCAN_PLANES) ? PM_SCAN_PLANES : PM_SCAN_AUTO;
    c->inner->stream = c->stream;
    for (size_t i = 0; i < ns; i++) {
      const std::string &ps = c->pats[i].s;
      const size_t l1 = ps.size() / 2;
      if (l1 == 0) return fail(PM_E_ARG, "exact_halves needs patterns of at least 2 characters");
      if ((rc = pm_add_pattern(c->inner, ps.substr(0, l1).c_str(), (uint32_t)(2 * i + 1), 0, 0))) return rc;
      if ((rc = pm_add_pattern(c->inner, ps.substr(l1).c_str(), (uint32_t)(2 * i + 2), 0, 0))) return rc;
    }
    if ((rc = pm_init(c->inner))) return rc;
  }
  // ---- pair-seed index (pm_index.cuh): applicable? ----------------------------------------
  if (!c->inner) {
    const int P = K + 2;
    const bool edit_mode = (K > 0 && c->opt.indels);
    // text codes: without wildcards code = symbol; with wildcards the codes are the match
    // classes of the text bytes 'A','C','G','T'
    const uint32_t symbits = (c->n_sym >= 32) ? 0xffffffffu : ((1u << c->n_sym) - 1u);
    uint32_t code_mask[4] = {0, 0, 0, 0};  // symbols matched by a text character of that code
    int n_codes = 0;
    bool ok = c->max_len <= 32 && ns > 0 && ns < (1u << 22) && K <= 3;   // k = 4, 5: the bit-plane kernels
    // DNA: code = (byte >> 1) & 3 for the bytes A, C, T, G (0, 1, 2, 3), which the pack kernel
    // computes without its table (pack_linear_kernel's fast path)
    bool dna_codes = wc;
    if (!wc) {
      if (c->n_sym > 4) ok = false;
      dna_codes = true;
      for (unsigned char ch : sym_char)
        if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') dna_codes = false;
      if (dna_codes) {
        n_codes = 4;
        for (int s2 = 0; s2 < c->n_sym; s2++) code_mask[(sym_char[s2] >> 1) & 3] = 1u << s2;
      } else {
        n_codes = std::min(c->n_sym, 4);
        for (int s2 = 0; s2 < n_codes; s2++) code_mask[s2] = 1u << s2;
      }
    } else {
      n_codes = 4;
      for (int cc = 0; cc < 4; cc++) code_mask[cc] = symmask[(unsigned char)"ACTG"[cc]] & symbits;
    }
    // every text byte must be one of the codes or match nothing
    std::vector<uint8_t> lin(256, 0);
    bool has_amb = false;
    for (int b2 = 0; b2 < 256 && ok; b2++) {
      const uint32_t msk = symmask[b2] & symbits;
      uint8_t v = 4;  // matches nothing
      if (msk) {
        int code = -1;
        for (int cc = 0; cc < n_codes; cc++)
          if (code_mask[cc] == msk) { code = cc; break; }
        if (code >= 0) v = (uint8_t)code;
        else if (wc) { v = 4 | 16; has_amb = true; }  // ambiguous: packed as "bad" and listed
        else ok = false;
      }
      // the four bytes of the fast path carry their arithmetic code even when no pattern
      // symbol matches them (no accept mask has that code set then: same as "matches nothing")
      if (dna_codes && (b2 == 'A' || b2 == 'C' || b2 == 'G' || b2 == 'T')) v = (uint8_t)((b2 >> 1) & 3);
      if (b2 == c->opt.eos) v = 4 | 8;
      lin[b2] = v;
    }
    c->idx_dna_fast = ok && dna_codes && c->opt.eos != 'A' && c->opt.eos != 'C' && c->opt.eos != 'G' &&
                      c->opt.eos != 'T';
    int min_len = 64;
    for (size_t i = 0; i < ns; i++) min_len = std::min<int>(min_len, (int)c->pats[i].s.size());
    int L = std::min(PAIR_LMAX, min_len / P);
    if (edit_mode) L = std::min(L, (32 - K) / P);
    if (L < 4) ok = false;
    if (c->scan_path == PM_SCAN_PLANES) ok = false;

    std::vector<uint2> ents;
    std::vector<uint32_t> dist;
    if (ok) {
      // edits: the distances some piece pair can have, (pj - pi) L + e with |e| <= K and never less
      // than L (for long pieces the ranges around L, 2L, .. leave gaps: those are not looked up)
      if (edit_mode) {
        for (int d = L; d <= (P - 1) * L + K; d++) {
          bool used = false;
          for (int g = 1; g < P && !used; g++) used = (d >= g * L - K && d <= g * L + K);
          if (used) dist.push_back((uint32_t)d);
        }
      } else {
        for (int j = 1; j < P; j++) dist.push_back((uint32_t)(j * L));
      }
      if ((int)dist.size() > PAIR_MAX_ND) ok = false;
    }
    auto dist_index = [&](int d) {
      for (size_t x = 0; x < dist.size(); x++)
        if ((int)dist[x] == d) return (uint32_t)x;
      return 0u;
    };
    // accept sets of a pattern character over the text codes
    auto accept4 = [&](unsigned char ch) {
      uint32_t acc = 0;
      const int sy = sym_of[ch];
      for (int cc = 0; cc < n_codes; cc++)
        if ((code_mask[cc] >> sy) & 1u) acc |= 1u << cc;
      return acc;
    };
    if (ok) {
      const size_t kMaxPerPair = 1024, kMaxEntries = (size_t)8 << 20;
      std::vector<uint32_t> ka, kb;
      auto expand = [&](const std::string &ps, int off, std::vector<uint32_t> &out) {
        out.assign(1, 0u);
        for (int t = 0; t < L; t++) {
          const uint32_t acc = accept4((unsigned char)ps[off + t]);
          std::vector<uint32_t> nxt;
          for (uint32_t base : out)
            for (int cc = 0; cc < 4; cc++)
              if ((acc >> cc) & 1u) nxt.push_back(base | ((uint32_t)cc << (2 * t)));
          out.swap(nxt);
          if (out.size() > kMaxPerPair) return false;
        }
        return true;
      };
      for (size_t i = 0; i < ns && ok; i++) {
        const std::string &ps = c->pats[i].s;
        std::vector<std::vector<uint32_t>> piece_keys(P);
        for (int pi = 0; pi < P && ok; pi++)
          if (!expand(ps, pi * L, piece_keys[pi])) ok = false;
        for (int pi = 0; pi < P && ok; pi++)
          for (int pj = pi + 1; pj < P && ok; pj++) {
            if (piece_keys[pi].size() * piece_keys[pj].size() > kMaxPerPair) { ok = false; break; }
            for (int e = edit_mode ? -K : 0; e <= (edit_mode ? K : 0); e++) {
              const int d = (pj - pi) * L + e;
              if (d < L) continue;
              const uint32_t di = dist_index(d);
              const uint32_t meta = ((uint32_t)i << 10) | (di << 5) | (uint32_t)(pi * L);
              for (uint32_t x : piece_keys[pi])
                for (uint32_t y : piece_keys[pj]) ents.push_back(make_uint2(x | (y << (2 * L)), meta));
            }
            if (ents.size() > kMaxEntries) ok = false;
          }
      }
    }
    if (c->scan_path == PM_SCAN_INDEX && !ok)
      return fail(PM_E_UNSUPPORTED, "pair-seed index scan not applicable to these patterns / options");
    c->use_index = c->index_ok = ok;
    c->idx_has_amb = ok && has_amb;
    c->idx_q = ok ? L : 0;
    if (ok) {
      IndexArgs &ia = c->idx_proto;
      memset(&ia, 0, sizeof ia);
      ia.log_words = PAIR_LOG_WORDS;
      ia.L = (uint32_t)L;
      ia.nd = (uint32_t)dist.size();
      ia.mult = 0x9E3779B1u;
      for (size_t d = 0; d < dist.size(); d++) {
        ia.dist[d] = dist[d];
        // substitutions: one salt per distance keeps the three key sets apart.  edits: every pair
        // is entered at 2k+1 neighbouring distances; sharing the bits between the distances
        // (salt 0) sets (2k+1) times fewer of them, and level 2 still tells the distances apart
        ia.salt[d] = edit_mode ? 0u : (uint32_t)d * 0x7F4A7C15u;
      }
      int lb = 10;
      while (((size_t)1 << lb) < 2 * ents.size() && lb < 26) lb++;
      ia.log_buckets = (uint32_t)lb;
      const size_t nb = (size_t)1 << lb;
      // level 1b: 64 KB of shared memory keyed by (key, distance) for edits; 16 MB of global
      // memory (L2-resident) once level 1 saturates (its 2^20 bits pass ~entries / 2^20 of the
      // positions)
      // measured crossover (1 Gbp, 2000-20000 primers): below ~100 k entries the queue stages cope with
      // level 1's false positives faster than 16 extra L2 loads per word do
      static const size_t big_set_entries = getenv("PM_BIG_SET_ENTRIES") ? (size_t)atoll(getenv("PM_BIG_SET_ENTRIES")) : 100000;
      // the edit kernel specialised on (L, k) reads level 1 through the bijective index when the keys fit
      static const bool edit_generic = getenv("PM_EDIT_GENERIC") && atoi(getenv("PM_EDIT_GENERIC")) != 0;
      c->idx_edit_special = edit_mode && !edit_generic && pair_edit_kernel(L, K) != nullptr;
      ia.l1_bij = (c->idx_edit_special && 4 * L <= 20 && ia.log_words == 15) ? 1u : 0u;
      // (that kernel asks level 1b from its queued stage, one L2 round trip per question if the table
      // were global: it keeps the shared-memory table until that is half full)
      const bool big_set = c->idx_edit_special ? ents.size() > ((size_t)1 << 18) : ents.size() > big_set_entries;
      ia.b2_global = big_set ? 1u : 0u;
      // substitutions, 24 k .. 100 k entries: level 1 passes 2-10 % of the lookups; the 64 KB
      // (key, distance) bitmap in the free shared memory removes most of them (+7-10 %); a kernel
      // variant of its own (pair_scan_sub_kernel<.., B1 = true>), the plain one has no register to spare
      static const size_t mid_set_entries = getenv("PM_MID_SET_ENTRIES") ? (size_t)atoll(getenv("PM_MID_SET_ENTRIES")) : 24000;
      const bool sub_1b = !edit_mode && !big_set && ents.size() > mid_set_entries;
      ia.log_words2 = big_set ? 22u : ((edit_mode || sub_1b) ? 14u : 0u);
      ia.mult2 = 0x85EBCA6Bu;
      for (size_t d = 0; d < dist.size(); d++) ia.salt2[d] = (uint32_t)(d + 1) * 0xC2B2AE35u;
      const bool have2 = ia.log_words2 != 0;
      std::vector<uint32_t> bitmap2(have2 ? ((size_t)1 << ia.log_words2) : 4, 0);
      std::vector<uint32_t> bitmap((size_t)1 << ia.log_words, 0), bucket(nb + 1, 0);
      std::vector<uint32_t> hs(ents.size());
      for (size_t x = 0; x < ents.size(); x++) {
        const uint32_t di = (ents[x].y >> 5) & 31u;
        if (ia.l1_bij) {
          // pair keys of at most 20 bits (L <= 5) against the 2^20 bits of level 1: an odd multiplier
          // mod 2^20 is a bijection, so a text key only ever meets the bit of the SAME key -- no
          // hash collisions (a non-injective hash costs 3x the false positives here: the key space
          // is no larger than the table, profiles/README.md)
          const uint32_t v = (ents[x].x * ia.mult) & 0xfffffu;
          bitmap[v >> 5] |= 1u << (31u - (ents[x].x & 31u));   // pair_probe_bij: bit chosen by the key itself
        } else {
          const uint32_t idx = pair_hash1(ents[x].x, ia.mult, ia.salt[di]);
          bitmap[idx >> (32 - ia.log_words)] |= 1u << (31u - (idx & 31u));
        }
        if (have2) {
          const uint32_t idx2 = ents[x].x * ia.mult2 + ia.salt2[di];
          bitmap2[idx2 >> (32 - ia.log_words2)] |= 1u << (idx2 & 31u);
        }
        hs[x] = pair_hash2(ents[x].x, di, ia.log_buckets);
        bucket[hs[x] + 1]++;
      }
      for (size_t b2 = 0; b2 < nb; b2++) bucket[b2 + 1] += bucket[b2];
      std::vector<uint32_t> fill(bucket.begin(), bucket.end() - 1);
      std::vector<uint2> sorted(std::max<size_t>(ents.size(), 1), make_uint2(0, 0));
      for (size_t x = 0; x < ents.size(); x++) sorted[fill[hs[x]]++] = ents[x];
      // expected seeds per base on uniform ACGT text (edit mode batches its seed list by this)
      c->idx_seed_rate = (double)ents.size() / std::pow(4.0, 2.0 * L);
      c->idx_entries = ents.size();
      // per-code accept masks (bit 2j) and Shift-And masks (bit j)
      std::vector<unsigned long long> pacc((ns + 1) * 4, 0);
      std::vector<uint32_t> um4((ns + 1) * 4, 0);
      for (size_t i = 0; i < ns; i++) {
        const std::string &ps = c->pats[i].s;
        for (int j = 0; j < (int)ps.size(); j++) {
          const uint32_t acc = accept4((unsigned char)ps[j]);
          for (int cc = 0; cc < 4; cc++)
            if ((acc >> cc) & 1u) {
              pacc[i * 4 + cc] |= 1ull << (2 * j);
              um4[i * 4 + cc] |= 1u << j;
            }
        }
      }
      {
        // plain patterns: every character accepts exactly one text code (the edit verifier then
        // compares codes with one XOR)
        bool plain = true;
        for (size_t i = 0; i < ns && plain; i++)
          for (int j = 0; j < (int)c->pats[i].s.size() && plain; j++) {
            int n_acc = 0;
            for (int cc = 0; cc < 4; cc++) n_acc += (int)((pacc[i * 4 + cc] >> (2 * j)) & 1ull);
            plain = (n_acc == 1);
          }
        ia.plain_patterns = plain ? 1u : 0u;
      }
      const Up iups[] = {{&c->d_bucket, bucket.data(), bucket.size() * 4},
                         {&c->d_entries, sorted.data(), sorted.size() * 8},
                         {&c->d_bitmap, bitmap.data(), bitmap.size() * 4},
                         {&c->d_bitmap2, bitmap2.data(), bitmap2.size() * 4},
                         {&c->d_pacc, pacc.data(), pacc.size() * 8},
                         {&c->d_umask4, um4.data(), um4.size() * 4},
                         {&c->d_lin, lin.data(), lin.size()}};
      for (const Up &u : iups) {
        int rc = u.b->ensure(u.bytes);
        if (rc) return rc;
        PM_CUDA(cudaMemcpy(u.b->p, u.src, u.bytes, cudaMemcpyHostToDevice));
      }
    }
  }
  c->inited = true;
  // a text set before init (not the reference's order, but harmless) is packed now
  if (c->d_text && c->n) {
    int rc = c->inner ? pm_set_text_device(c->inner, c->d_text, c->n, c->offset) : pack_text(c);
    if (rc) return rc;
    PM_CUDA(cudaStreamSynchronize(c->stream));
  }
  return PM_OK;
}

// exact_halves: the text lives in (and is packed for) the inner context; this one adopts it
static int adopt_inner_text(pm_ctx *c, int rc) {
  if (rc) return rc;
  c->d_text = c->inner->d_text;
  c->n = c->inner->n;
  c->offset = c->inner->offset;
  c->seq_key = c->inner->seq_key;
  c->seq_hdr_off = c->inner->seq_hdr_off;
  c->seq_hdr_len = c->inner->seq_hdr_len;
  return PM_OK;
}

static int set_text_common(pm_ctx *c, uint64_t n, int64_t offset) {
  c->use_index = c->index_ok;
  c->n = n;
  c->offset = offset;
  if (n >= (1ull << 40)) return fail(PM_E_UNSUPPORTED, "shard longer than 2^40 characters");
  if (!c->inited) return PM_OK;  // packed by pm_init
  return pack_text(c);
}

// Host text -> device, as a pipeline on the private copy stream: the source moves in
// growing chunks; every chunk is (for a .sqz source) unpacked to stream bytes and then
// packed tile by tile as far as its characters (plus the 64-character extension) reach;
// the scan of pm_find_patterns waits chunk by chunk, so it overlaps the rest of the
// transfer.  bits == 0: the source is the stream itself.
static int unpack_groups(pm_ctx *c, uint32_t bits, uint64_t n_bytes, uint64_t g0, uint64_t g1,
                         cudaStream_t st) {
  if (g1 <= g0) return PM_OK;
  const uint64_t per_block = 256ull * SQZ_GROUPS_PER_THREAD;
  const unsigned blocks = (unsigned)((g1 - g0 + per_block - 1) / per_block);
  const uint8_t *src = c->d_sqz.as<uint8_t>(), *tab = c->d_sqztab.as<uint8_t>();
  uint8_t *dst = c->d_text_own.as<uint8_t>();
#define PM_UNPACK(B)                                                                            \
  case B:                                                                                       \
    unpack_sqz_kernel<B><<<blocks, 256, 0, st>>>(src, n_bytes, tab, c->opt.eos, dst, g0, g1, c->n); \
    break;
  switch (bits) {
    PM_UNPACK(1) PM_UNPACK(2) PM_UNPACK(3) PM_UNPACK(4) PM_UNPACK(5) PM_UNPACK(6) PM_UNPACK(7) PM_UNPACK(8)
    default: return fail(PM_E_ARG, "bits per character out of range");
  }
#undef PM_UNPACK
  PM_CUDA(cudaGetLastError());
  return PM_OK;
}

// src: the bytes in host memory, or (src == nullptr) `reader`, which delivers them chunk by chunk
// from a file through its ring of page-locked buffers (chunk size = this pipeline's)
static uint64_t stream_chunk_bytes(uint32_t bits) { return bits ? (24ull << 20) : (64ull << 20); }

static int stream_in(pm_ctx *c, const uint8_t *src, uint64_t n_bytes, uint32_t bits, uint64_t n,
                     int64_t offset, pm::ChunkReader *reader = nullptr) {
  int rc = c->d_text_own.ensure(n + 64);
  if (rc) return rc;
  c->d_text = c->d_text_own.as<uint8_t>();
  c->use_index = c->index_ok;
  c->n = n;
  c->offset = offset;
  if (!n) return set_text_common(c, n, offset);
  uint8_t *d_dst = c->d_text_own.as<uint8_t>();
  if (bits) {
    if ((rc = c->d_sqz.ensure(n_bytes + 64))) return rc;
    d_dst = c->d_sqz.as<uint8_t>();
  }
  cudaPointerAttributes at;
  bool pinned = false;
  if (reader) pinned = true;
  else if (cudaPointerGetAttributes(&at, src) == cudaSuccess)
    pinned = (at.type == cudaMemoryTypeHost);
  else
    cudaGetLastError();
  if (!c->inited && reader) return fail(PM_E_ARG, "pm_set_text_database before pm_init");
  if (!c->inited) {  // not the reference's call order: plain copy now, pack in pm_init
    PM_CUDA(cudaMemcpyAsync(d_dst, src, n_bytes, cudaMemcpyHostToDevice, c->stream));
    if (bits && (rc = unpack_groups(c, bits, n_bytes, 0, (n + 15) / 16, c->stream))) return rc;
    PM_CUDA(cudaStreamSynchronize(c->stream));
    return PM_OK;
  }
  if ((rc = prepare_planes(c))) return rc;
  PM_CUDA(cudaEventRecord(c->ev[4], c->stream));
  PM_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev[4], 0));  // earlier work may still read the buffers
  PM_CUDA(cudaStreamWaitEvent(c->prep_stream, c->ev[4], 0));
  PM_CUDA(cudaMemsetAsync(c->d_counters + 8, 0, sizeof(unsigned long long), c->prep_stream));
  if (!pinned)
    for (int i = 0; i < 2; i++)
      if (!c->stage[i]) PM_CUDA(cudaMallocHost(&c->stage[i], kStageBytes));
  // constant chunks: what is left to do when the last byte has arrived is one chunk's
  // unpack + pack + scan, so small chunks keep the tail short (a chunk costs ~6 launches)
  const uint64_t chunk_bytes = stream_chunk_bytes(bits);
  uint64_t done = 0, groups_done = 0;
  uint32_t tiles_done = 0;
  int which = 0;
  while (done < n_bytes) {
    uint64_t take = std::min<uint64_t>(chunk_bytes, n_bytes - done);
    if (!reader && n_bytes - done - take < chunk_bytes / 4) take = n_bytes - done;  // no tiny last chunk
    if (reader) {
      uint64_t got = 0;
      const uint8_t *p = reader->acquire(done / chunk_bytes, &got);
      if (!p || got < take) return fail(PM_E_ARG, "read error on the database file");
      PM_CUDA(cudaMemcpyAsync(d_dst + done, p, take, cudaMemcpyHostToDevice, c->copy_stream));
      reader->release(done / chunk_bytes, c->copy_stream);
    } else if (pinned) {
      PM_CUDA(cudaMemcpyAsync(d_dst + done, src + done, take, cudaMemcpyHostToDevice, c->copy_stream));
    } else {
      for (uint64_t o = 0; o < take; o += kStageBytes) {
        const size_t part = (size_t)std::min<uint64_t>(kStageBytes, take - o);
        PM_CUDA(cudaEventSynchronize(c->stage_ev[which]));
        memcpy(c->stage[which], src + done + o, part);
        PM_CUDA(cudaMemcpyAsync(d_dst + done + o, c->stage[which], part, cudaMemcpyHostToDevice,
                                c->copy_stream));
        PM_CUDA(cudaEventRecord(c->stage_ev[which], c->copy_stream));
        which ^= 1;
      }
    }
    done += take;
    const bool last = (done == n_bytes);
    uint64_t chars = n;  // characters of the stream that are on the device now
    if (!last) chars = bits ? std::min<uint64_t>(n, (done * 8 / bits) & ~15ull) : done;
    // the kernels of this chunk run on the prep stream, so that the next copy starts at once
    cudaEvent_t arrived = next_event(c);
    PM_CUDA(cudaEventRecord(arrived, c->copy_stream));
    PM_CUDA(cudaStreamWaitEvent(c->prep_stream, arrived, 0));
    if (bits) {
      const uint64_t g1 = last ? (n + 15) / 16 : chars / 16;
      if ((rc = unpack_groups(c, bits, n_bytes, groups_done, g1, c->prep_stream))) return rc;
      groups_done = std::max(groups_done, g1);
    }
    const uint32_t ready = last ? c->n_tiles
                                : (uint32_t)((chars > (uint64_t)TILE_EXT ? chars - TILE_EXT : 0) / TILE_POS);
    if ((rc = pack_tiles(c, tiles_done, ready, c->prep_stream))) return rc;
    cudaEvent_t e = next_event(c);
    PM_CUDA(cudaEventRecord(e, c->prep_stream));
    c->chunks.push_back(pm_ctx::Chunk{ready, e});
    tiles_done = std::max(tiles_done, ready);
  }
  return PM_OK;
}

int pm_set_text_host(pm_ctx *c, const uint8_t *stream, uint64_t n, int64_t offset) {
  if (!c || (!stream && n)) return fail(PM_E_ARG, "null argument");
  if (c->inner) return adopt_inner_text(c, pm_set_text_host(c->inner, stream, n, offset));
  PM_CUDA(cudaSetDevice(c->opt.device));
  if (n && offset == 0 && stream[0] != c->opt.eos)
    return fail(PM_E_ARG, "a stream at offset 0 must begin with the end-of-sequence character");
  if (n >= (1ull << 40)) return fail(PM_E_UNSUPPORTED, "shard longer than 2^40 characters");
  return stream_in(c, stream, n, 0, n, offset);
}

int pm_sqz_bits(uint32_t table_size) {  // char_io.t:58-61
  if (table_size == 0 || table_size > 256) return fail(PM_E_ARG, "character table size must be 1..256");
  int bits = 1;
  while ((1u << bits) < table_size) bits++;
  return bits;
}

static int set_text_coded(pm_ctx *c, const uint8_t *sqz, uint64_t n_bytes, const uint8_t *table,
                          uint32_t table_size, uint64_t n_chars, int64_t offset, int bits);

int pm_set_text_sqz_host(pm_ctx *c, const uint8_t *sqz, uint64_t n_bytes, const uint8_t *table,
                         uint32_t table_size, uint64_t n_chars, int64_t offset) {
  if (!c || (!sqz && n_bytes) || !table) return fail(PM_E_ARG, "null argument");
  if (c->inner) return adopt_inner_text(c, pm_set_text_sqz_host(c->inner, sqz, n_bytes, table, table_size, n_chars, offset));
  const int bits = pm_sqz_bits(table_size);
  if (bits < 0) return bits;
  return set_text_coded(c, sqz, n_bytes, table, table_size, n_chars, offset, bits);
}

// Normalized<> (char_io.t:216-283): the .sqn file of compress_seq -n, one table index per byte
int pm_set_text_sqn_host(pm_ctx *c, const uint8_t *sqn, uint64_t n_chars, const uint8_t *table,
                         uint32_t table_size, int64_t offset) {
  if (!c || (!sqn && n_chars) || !table) return fail(PM_E_ARG, "null argument");
  if (table_size == 0 || table_size > 256) return fail(PM_E_ARG, "character table size must be 1..256");
  if (c->inner) return adopt_inner_text(c, pm_set_text_sqn_host(c->inner, sqn, n_chars, table, table_size, offset));
  return set_text_coded(c, sqn, n_chars, table, table_size, n_chars, offset, 8);
}

static int set_text_coded(pm_ctx *c, const uint8_t *sqz, uint64_t n_bytes, const uint8_t *table,
                          uint32_t table_size, uint64_t n_chars, int64_t offset, int bits) {
  PM_CUDA(cudaSetDevice(c->opt.device));
  const uint64_t held = n_bytes * 8 / (uint64_t)bits;
  if (n_chars == 0) n_chars = held;  // the writer's padding reads as end-of-sequence characters
  if (n_chars > held) return fail(PM_E_ARG, "n_chars exceeds what the compressed bytes hold");
  if (n_chars >= (1ull << 40)) return fail(PM_E_UNSUPPORTED, "shard longer than 2^40 characters");
  if (n_chars && offset == 0 && table[(sqz[0] >> (8 - bits)) & ((1u << bits) - 1u)] != c->opt.eos)
    return fail(PM_E_ARG, "a stream at offset 0 must begin with the end-of-sequence character");
  uint8_t tab[256];
  memset(tab, 0, sizeof tab);
  memcpy(tab, table, table_size);
  int rc = c->d_sqztab.ensure(256);
  if (rc) return rc;
  // the table is read by kernels on the copy stream, which starts after ev[4] of c->stream
  PM_CUDA(cudaMemcpyAsync(c->d_sqztab.p, tab, 256, cudaMemcpyHostToDevice, c->stream));
  PM_CUDA(cudaStreamSynchronize(c->stream));
  return stream_in(c, sqz, n_bytes, (uint32_t)bits, n_chars, offset);
}

int pm_set_text_device(pm_ctx *c, const uint8_t *d_stream, uint64_t n, int64_t offset) {
  if (!c || (!d_stream && n)) return fail(PM_E_ARG, "null argument");
  if (c->inner) return adopt_inner_text(c, pm_set_text_device(c->inner, d_stream, n, offset));
  if ((reinterpret_cast<uintptr_t>(d_stream) & 3) != 0)
    return fail(PM_E_ARG, "device stream must be 4-byte aligned");
  PM_CUDA(cudaSetDevice(c->opt.device));
  c->d_text = d_stream;
  return set_text_common(c, n, offset);
}

int pm_shard_overlap(pm_ctx *c) {
  if (!c) return fail(PM_E_ARG, "null context");
  return c->max_len + 3 * c->opt.k + 2;
}

int pm_find_patterns_device(pm_ctx *c, int64_t begin, int64_t end, const pm_hit **d_hits,
                            uint64_t *nhits) {
  if (!c || !d_hits || !nhits) return fail(PM_E_ARG, "null argument");
  *d_hits = nullptr;
  int rc = find_patterns_impl(c, begin, end, nhits, false);
  if (*nhits) *d_hits = c->d_hits2.as<pm_hit>();
  return rc;
}

int pm_find_patterns(pm_ctx *c, int64_t begin, int64_t end, const pm_hit **hits, uint64_t *nhits) {
  if (!c || !hits || !nhits) return fail(PM_E_ARG, "null argument");
  *hits = nullptr;
  int rc = find_patterns_impl(c, begin, end, nhits, true);   // records land in c->h_out
  if (*nhits) *hits = reinterpret_cast<const pm_hit *>(c->h_out);
  return rc;
}

int pm_scan_candidates(pm_ctx *c, int64_t begin, int64_t end, const pm_candidate **cands,
                       uint64_t *ncands) {
  if (!c || !cands || !ncands) return fail(PM_E_ARG, "null argument");
  *cands = nullptr;
  *ncands = 0;
  if (!c->inited) return fail(PM_E_ARG, "pm_scan_candidates before pm_init");
  if (c->inner)
    return fail(PM_E_UNSUPPORTED, "raw candidates are defined for the bit-vector and exact indexes, not exact_halves");
  PM_CUDA(cudaSetDevice(c->opt.device));
  memset(&c->stats, 0, offsetof(pm_stats, pack_ms));
  c->stats.text_len = c->n;
  c->stats.n_patterns = c->pats.size();
  int64_t lo = begin - c->offset, hi = end - c->offset;
  if (begin < c->offset || lo < 0) lo = 0;
  if (end == std::numeric_limits<int64_t>::max() || hi > (int64_t)c->n + 1) hi = (int64_t)c->n + 1;
  uint64_t n = 0;
  int rc = run_scan_any(c, lo, hi, &n, true);
  if (rc) return rc;
  c->stats.n_candidates = n;
  if (!n) return PM_OK;
  if ((rc = c->d_cands.ensure(n * sizeof(pm_candidate)))) return rc;
  keys_to_cands_kernel<<<nblocks(n, 256), 256, 0, c->stream>>>(
      c->d_keys.as<unsigned long long>(), n, c->d_slotid.as<uint32_t>(), c->offset,
      c->d_cands.as<pm_candidate>());
  PM_CUDA(cudaGetLastError());
  if ((rc = ensure_host_out(c, n * sizeof(pm_candidate)))) return rc;
  PM_CUDA(cudaMemcpyAsync(c->h_out, c->d_cands.p, n * sizeof(pm_candidate), cudaMemcpyDeviceToHost,
                          c->stream));
  PM_CUDA(cudaStreamSynchronize(c->stream));
  *cands = reinterpret_cast<const pm_candidate *>(c->h_out);
  *ncands = n;
  return PM_OK;
}

int pm_last_candidates(pm_ctx *c, const pm_candidate **cands, uint64_t *ncands) {
  if (!c || !cands || !ncands) return fail(PM_E_ARG, "null argument");
  *cands = nullptr;
  *ncands = 0;
  if (!c->inited) return fail(PM_E_ARG, "pm_last_candidates before pm_init");
  if (c->inner)
    return fail(PM_E_UNSUPPORTED, "the hits of exact_halves / exact_bases do not come from a candidate list");
  const uint64_t n = c->last_uniq;
  if (!n) return PM_OK;
  PM_CUDA(cudaSetDevice(c->opt.device));
  int rc;
  if ((rc = c->d_cands.ensure(n * sizeof(pm_candidate)))) return rc;
  keys_to_cands_kernel<<<nblocks(n, 256), 256, 0, c->stream>>>(
      c->d_keys.as<unsigned long long>(), n, c->d_slotid.as<uint32_t>(), c->offset,
      c->d_cands.as<pm_candidate>());
  PM_CUDA(cudaGetLastError());
  c->h_cands.resize(n);
  PM_CUDA(cudaMemcpyAsync(c->h_cands.data(), c->d_cands.p, n * sizeof(pm_candidate), cudaMemcpyDeviceToHost,
                          c->stream));
  PM_CUDA(cudaStreamSynchronize(c->stream));
  *cands = c->h_cands.data();
  *ncands = n;
  return PM_OK;
}

int pm_hit_order(const pm_hit *hits, uint64_t nhits, const pm_candidate *cands, uint64_t ncands, int k,
                 uint64_t *order) {
  if ((nhits && (!hits || !order)) || (ncands && !cands)) return fail(PM_E_ARG, "null argument");
  std::vector<uint64_t> o;
  hit_order(hits, nhits, cands, ncands, k, o);
  for (uint64_t i = 0; i < nhits; i++) order[i] = o[i];
  return PM_OK;
}

int pm_reset(pm_ctx *c) {
  if (!c) return fail(PM_E_ARG, "null context");
  if (c->inner) pm_reset(c->inner);
  c->d_text = nullptr;
  c->n = 0;
  c->n_tiles = 0;
  return PM_OK;
}

int pm_get_stats(pm_ctx *c, pm_stats *out) {
  if (!c || !out) return fail(PM_E_ARG, "null argument");
  *out = c->stats;
  return PM_OK;
}

// StreamedFastaFile::getch, fasta_io.t:478-629 (eos_start = true): the stream
// is eos, seq1, eos, seq2, ..., seqN, eos.  Sequence lines are joined; blanks,
// CR and LF inside sequence data are dropped; '>' starts a header anywhere.
// stop_early: return as soon as the first stream character after the initial eos
// exists (*rp_out = raw bytes consumed); the closing eos is then not written.
static void normalize_walk(const uint8_t *raw, uint64_t n, int upper_case, unsigned char eos, uint8_t *out,
                           uint64_t *out_len, int64_t *seq_key, int64_t *hdr_off, int64_t *hdr_len,
                           uint64_t max_seqs, uint64_t *n_seqs, bool stop_early, uint64_t *rp_out) {
  uint64_t rp = 0, pos = 0, nseq = 0;
  int64_t lastkey = 0;
  auto put = [&](uint8_t ch) {
    if (upper_case && ch >= 'a' && ch <= 'z') ch = (uint8_t)(ch - 'a' + 'A');
    out[pos++] = ch;
  };
  if (n > 0) out[pos++] = eos;  // the stream opens with an end-of-sequence character
  while (rp < n) {
    // (and the raw position is 16-byte aligned, so that the device can load uint4; a header
    // consumes many bytes at once, hence the cap)
    if (stop_early && pos > 1 && ((rp & 15) == 0 || rp > 4096)) break;
    uint8_t ch = raw[rp++];
    const bool blank = (ch == '\n' || ch == '\r' || ch == ' ');
    if (!blank && ch != '>') {
      put(ch);
      continue;
    }
    bool at_eof = false;
    while (ch == '\n' || ch == '\r' || ch == ' ') {
      if (rp >= n) {
        at_eof = true;
        break;
      }
      ch = raw[rp++];
    }
    if (at_eof) break;
    if (ch != '>') {
      put(ch);
      continue;
    }
    // header line
    const uint64_t hstart = rp;
    uint8_t term = 0;
    while (rp < n) {
      term = raw[rp++];
      if (term == '\n' || term == '\r') break;
    }
    const uint64_t hend = rp;
    if (term == '\r' && rp < n) rp++;
    const bool first = (pos == 1);
    if (!first) pos++;  // room for the separator
    if (lastkey < (int64_t)pos || first) {
      if (nseq < max_seqs) {
        if (seq_key) seq_key[nseq] = (int64_t)pos;
        if (hdr_off) hdr_off[nseq] = (int64_t)hstart;
        if (hdr_len) hdr_len[nseq] = (int64_t)(hend - hstart) - 1;
      }
      nseq++;
      lastkey = (int64_t)pos;
    }
    if (first) {
      if (rp < n) put(raw[rp++]);  // the character after the first header is taken verbatim
    } else {
      out[pos - 1] = eos;
    }
  }
  if (!(stop_early && rp < n)) out[pos++] = eos;
  *out_len = pos;
  if (n_seqs) *n_seqs = nseq;
  if (rp_out) *rp_out = rp;
}

int pm_normalize_fasta(const uint8_t *raw, uint64_t n, int upper_case, unsigned char eos,
                       uint8_t *out, uint64_t *out_len, int64_t *seq_key, int64_t *hdr_off,
                       int64_t *hdr_len, uint64_t max_seqs, uint64_t *n_seqs) {
  if ((!raw && n) || !out || !out_len) return fail(PM_E_ARG, "null argument");
  normalize_walk(raw, n, upper_case, eos, out, out_len, seq_key, hdr_off, hdr_len, max_seqs, n_seqs, false,
                 nullptr);
  return PM_OK;
}

// The same on the device (pm_fasta.cuh): raw FASTA bytes in host memory -> stream in HBM.
constexpr uint64_t kFastaChunk = (uint64_t)FA_THREADS * FA_SEG * 2048;   // 64 MiB of raw bytes, whole blocks

// raw: the bytes of the file in host memory, or (raw == nullptr) `reader`, which delivers them in
// chunks of kFastaChunk through its ring of page-locked buffers
static int set_text_fasta_impl(pm_ctx *c, const uint8_t *raw, uint64_t n_raw, int upper_case, pm::ChunkReader *reader) {
  if (n_raw >= (1ull << 40)) return fail(PM_E_UNSUPPORTED, "file longer than 2^40 bytes");
  PM_CUDA(cudaSetDevice(c->opt.device));
  c->hdr_ready = false;
  c->hdr_from_file = false;
  const unsigned char eos = c->opt.eos;
  std::vector<uint8_t> whole;   // reader + a file that has to take the host path
  if (reader && n_raw) {
    uint64_t got = 0;
    const uint8_t *p0 = reader->acquire(0, &got);
    if (!p0) return fail(PM_E_ARG, "read error on the database file");
    raw = p0;   // the first chunk: the prefix walk below reads at most 16 KB of it
    if (n_raw < 16384) {
      whole.assign(p0, p0 + n_raw);
      reader->release(0, c->copy_stream);
      raw = whole.data();
      reader = nullptr;
    }
  }
  // host prefix: everything up to the first stream character (first-header rule)
  std::vector<uint8_t> pre;
  std::vector<int64_t> pk, po, pl;
  uint64_t pre_len = 0, pre_seqs = 0, r0 = 0;
  {
    // the early-stopping walk reads at most ~4 KB past the first stream character
    const bool tiny = n_raw < 16384;
    const size_t room = tiny ? n_raw + 4 : 16384 + 4;
    std::vector<uint8_t> tmp(room);
    pk.assign(room / 2 + 4, 0);
    po.assign(room / 2 + 4, 0);
    pl.assign(room / 2 + 4, 0);
    bool all = tiny;
    if (!all) {
      // find how far the walk would go without writing past the buffer: the stream never grows
      // faster than the raw position, and the walk stops within 4096 + one header of raw bytes;
      // a header longer than the buffer sends the whole file down the host path
      normalize_walk(raw, std::min<uint64_t>(n_raw, 16384), upper_case, eos, tmp.data(), &pre_len, pk.data(),
                     po.data(), pl.data(), pk.size(), &pre_seqs, true, &r0);
      if (r0 >= std::min<uint64_t>(n_raw, 16384)) all = true;  // no stream character in the first 16 KB
    }
    if (all && reader) {
      // (no stream character in the first 16 KB: the whole file through host memory)
      whole.resize(n_raw);
      for (uint64_t j = 0; j < reader->n_chunks(); j++) {
        uint64_t got = 0;
        const uint8_t *pj = reader->acquire(j, &got);
        if (!pj) return fail(PM_E_ARG, "read error on the database file");
        memcpy(whole.data() + j * reader->chunk_bytes(), pj, got);
        reader->release(j, c->copy_stream);
      }
      raw = whole.data();
      reader = nullptr;
    }
    if (all) {
      std::vector<uint8_t> full(n_raw + 4);
      std::vector<int64_t> k2(n_raw / 2 + 4), o2(n_raw / 2 + 4), l2(n_raw / 2 + 4);
      uint64_t len = 0, ns = 0;
      normalize_walk(raw, n_raw, upper_case, eos, full.data(), &len, k2.data(), o2.data(), l2.data(),
                     k2.size(), &ns, false, nullptr);
      c->seq_key.assign(k2.begin(), k2.begin() + ns);
      c->seq_hdr_off.assign(o2.begin(), o2.begin() + ns);
      c->seq_hdr_len.assign(l2.begin(), l2.begin() + ns);
      // the raw bytes go to the device as well: pm_get_header gathers the header lines from there
      if (int rc2 = c->d_fa_raw.ensure(n_raw + 64)) return rc2;
      PM_CUDA(cudaMemcpyAsync(c->d_fa_raw.p, raw, n_raw, cudaMemcpyHostToDevice, c->stream));
      PM_CUDA(cudaStreamSynchronize(c->stream));
      return stream_in(c, full.data(), len, 0, len, 0);
    }
    pre.assign(tmp.begin(), tmp.begin() + pre_len);
  }
  // The file is taken in chunks: the copy stream moves chunk j+1 while the prep stream
  // normalises chunk j (summary -> scan of the state maps -> counts -> two exclusive sums ->
  // write); the state / stream length / header count after a chunk are carried on the device
  // (d_counters[11..13]), so the host does not wait between chunks.
  const uint64_t kChunk = kFastaChunk;
  const uint64_t n_seg_max = std::min<uint64_t>(kChunk, n_raw - r0 + FA_SEG) / FA_SEG + 1;
  const uint64_t hdr_cap = std::min<uint64_t>(n_raw / 32 + 1024, 1ull << 27);
  int rc;
  if ((rc = c->d_fa_raw.ensure(n_raw + 64))) return rc;
  if ((rc = c->d_text_own.ensure(n_raw + 2 + 64))) return rc;
  if ((rc = c->d_fa_map.ensure(n_seg_max))) return rc;
  if ((rc = c->d_fa_scan.ensure(n_seg_max))) return rc;
  if ((rc = c->d_fa_cnt.ensure(n_seg_max * 8))) return rc;
  if ((rc = c->d_fa_chdr.ensure(n_seg_max * 4))) return rc;
  if ((rc = c->d_fa_start.ensure(n_seg_max))) return rc;
  if ((rc = c->d_fa_nout.ensure((n_seg_max + 1) * 4))) return rc;
  if ((rc = c->d_fa_nhdr.ensure((n_seg_max + 1) * 4))) return rc;
  if ((rc = c->d_fa_ooff.ensure((n_seg_max + 1) * 8))) return rc;
  if ((rc = c->d_fa_hoff.ensure((n_seg_max + 1) * 8))) return rc;
  if ((rc = c->d_fa_key.ensure(hdr_cap * 8))) return rc;
  if ((rc = c->d_fa_hs.ensure(hdr_cap * 8))) return rc;
  if ((rc = c->d_fa_he.ensure(hdr_cap * 8))) return rc;
  {
    size_t tb = 0, tb2 = 0;
    cub::DeviceScan::InclusiveScan(nullptr, tb, c->d_fa_map.as<uint8_t>(), c->d_fa_scan.as<uint8_t>(), FaCompose(),
                                   (int64_t)n_seg_max, c->prep_stream);
    cub::DeviceScan::ExclusiveScan(nullptr, tb2, c->d_fa_nout.as<uint32_t>(), c->d_fa_ooff.as<unsigned long long>(),
                                   cub::Sum(), 0ull, (int64_t)(n_seg_max + 1), c->prep_stream);
    if ((rc = c->d_fa_tmp.ensure(std::max(tb, tb2)))) return rc;
  }
  PM_CUDA(cudaFuncSetAttribute(fasta_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_WRITE_SMEM));
  cudaStream_t cs = c->copy_stream, st = c->prep_stream;
  unsigned long long *carry = c->d_counters + 11;   // [state, stream bytes, headers], + [14] overflow
  PM_CUDA(cudaEventRecord(c->ev[4], c->stream));    // earlier work of the context may still read the buffers
  PM_CUDA(cudaStreamWaitEvent(cs, c->ev[4], 0));
  PM_CUDA(cudaStreamWaitEvent(st, c->ev[4], 0));
  PM_CUDA(cudaMemsetAsync(carry, 0, 4 * sizeof(unsigned long long), st));
  fill_i64_kernel<<<nblocks(hdr_cap, 256), 256, 0, st>>>(c->d_fa_he.as<long long>(), hdr_cap, (long long)n_raw);
  PM_CUDA(cudaGetLastError());
  cudaPointerAttributes pat_;
  bool pinned = false;
  if (reader) pinned = true;
  else if (cudaPointerGetAttributes(&pat_, raw) == cudaSuccess) pinned = (pat_.type == cudaMemoryTypeHost);
  else cudaGetLastError();
  if (!pinned)
    for (int i = 0; i < 2; i++)
      if (!c->stage[i]) PM_CUDA(cudaMallocHost(&c->stage[i], kStageBytes));
  int which = 0;
  uint8_t *d_raw_w = c->d_fa_raw.as<uint8_t>();
  const uint8_t *d_raw = d_raw_w;
  uint8_t *d_stream = c->d_text_own.as<uint8_t>();
  c->ev_used = 0;
  // the bytes in front of the first chunk (the first header line lives there; pm_get_header reads it)
  if (r0) PM_CUDA(cudaMemcpyAsync(d_raw_w, raw, r0, cudaMemcpyHostToDevice, cs));
  // chunks end on multiples of kChunk (the reader's chunks); the first one starts behind the prefix
  for (uint64_t a0 = r0, a1 = 0; a0 < n_raw; a0 = a1) {
    a1 = std::min(n_raw, (a0 / kChunk + 1) * kChunk);
    if (reader) {
      const uint64_t j = a0 / kChunk;
      uint64_t got = 0;
      const uint8_t *pj = reader->acquire(j, &got);
      if (!pj || j * kChunk + got < a1) return fail(PM_E_ARG, "read error on the database file");
      PM_CUDA(cudaMemcpyAsync(d_raw_w + a0, pj + (a0 - j * kChunk), a1 - a0, cudaMemcpyHostToDevice, cs));
      reader->release(j, cs);
    } else if (pinned) {
      PM_CUDA(cudaMemcpyAsync(d_raw_w + a0, raw + a0, a1 - a0, cudaMemcpyHostToDevice, cs));
    } else {
      for (uint64_t o = a0; o < a1; o += kStageBytes) {
        const size_t part = (size_t)std::min<uint64_t>(kStageBytes, a1 - o);
        PM_CUDA(cudaEventSynchronize(c->stage_ev[which]));
        memcpy(c->stage[which], raw + o, part);
        PM_CUDA(cudaMemcpyAsync(d_raw_w + o, c->stage[which], part, cudaMemcpyHostToDevice, cs));
        PM_CUDA(cudaEventRecord(c->stage_ev[which], cs));
        which ^= 1;
      }
    }
    cudaEvent_t arrived = next_event(c);
    PM_CUDA(cudaEventRecord(arrived, cs));
    PM_CUDA(cudaStreamWaitEvent(st, arrived, 0));
    const uint64_t n_seg = (a1 - a0 + FA_SEG - 1) / FA_SEG;
    const unsigned fb = (unsigned)((n_seg + FA_THREADS - 1) / FA_THREADS);
    // the kernels index raw bytes from the chunk start; n_raw bounds the last chunk
    fasta_summary_kernel<<<fb, FA_THREADS, 0, st>>>(d_raw, a0, a1, n_seg, c->d_fa_map.as<uint8_t>(),
                                                    c->d_fa_cnt.as<unsigned long long>(), c->d_fa_chdr.as<uint32_t>());
    size_t tb = c->d_fa_tmp.cap;
    PM_CUDA(cub::DeviceScan::InclusiveScan(c->d_fa_tmp.p, tb, c->d_fa_map.as<uint8_t>(), c->d_fa_scan.as<uint8_t>(),
                                           FaCompose(), (int64_t)n_seg, st));
    PM_CUDA(cudaMemsetAsync(c->d_fa_nout.as<uint32_t>() + n_seg, 0, 4, st));
    PM_CUDA(cudaMemsetAsync(c->d_fa_nhdr.as<uint32_t>() + n_seg, 0, 4, st));
    fasta_select_kernel<<<nblocks(n_seg, 256), 256, 0, st>>>(
        c->d_fa_scan.as<uint8_t>(), c->d_fa_cnt.as<unsigned long long>(), c->d_fa_chdr.as<uint32_t>(), n_seg,
        c->d_fa_start.as<uint8_t>(), c->d_fa_nout.as<uint32_t>(), c->d_fa_nhdr.as<uint32_t>(), carry);
    tb = c->d_fa_tmp.cap;
    PM_CUDA(cub::DeviceScan::ExclusiveScan(c->d_fa_tmp.p, tb, c->d_fa_nout.as<uint32_t>(),
                                           c->d_fa_ooff.as<unsigned long long>(), cub::Sum(), 0ull,
                                           (int64_t)(n_seg + 1), st));
    tb = c->d_fa_tmp.cap;
    PM_CUDA(cub::DeviceScan::ExclusiveScan(c->d_fa_tmp.p, tb, c->d_fa_nhdr.as<uint32_t>(),
                                           c->d_fa_hoff.as<unsigned long long>(), cub::Sum(), 0ull,
                                           (int64_t)(n_seg + 1), st));
    fasta_write_kernel<<<fb, FA_THREADS, FA_WRITE_SMEM, st>>>(
        d_raw, a0, a1, n_seg, c->d_fa_start.as<uint8_t>(), c->d_fa_ooff.as<unsigned long long>(),
        c->d_fa_hoff.as<unsigned long long>(), upper_case, eos, d_stream, pre_len, pre_seqs,
        c->d_fa_key.as<long long>(), c->d_fa_hs.as<long long>(), c->d_fa_he.as<long long>(), carry, hdr_cap,
        carry + 3);
    fasta_carry_kernel<<<1, 32, 0, st>>>(c->d_fa_scan.as<uint8_t>(), n_seg, c->d_fa_ooff.as<unsigned long long>(),
                                         c->d_fa_hoff.as<unsigned long long>(), carry);
    PM_CUDA(cudaGetLastError());
  }
  unsigned long long totals[4] = {0, 0, 0, 0};
  PM_CUDA(cudaMemcpyAsync(totals, carry, sizeof totals, cudaMemcpyDeviceToHost, st));
  PM_CUDA(cudaStreamSynchronize(st));
  if (totals[3]) return fail(PM_E_UNSUPPORTED, "more FASTA records than the sequence table holds (2^27)");
  const uint64_t n = pre_len + totals[1] + 1;       // + the closing eos
  const uint64_t n_hdr = pre_seqs + totals[2];
  if (n > n_raw + 2) return fail(PM_E_CUDA, "FASTA normalisation produced an impossible length");
  PM_CUDA(cudaMemcpyAsync(d_stream, pre.data(), pre_len, cudaMemcpyHostToDevice, st));
  PM_CUDA(cudaMemsetAsync(d_stream + (n - 1), eos, 1, st));
  // sequence table back to the host
  c->seq_key.assign(n_hdr, 0);
  c->seq_hdr_off.assign(n_hdr, 0);
  c->seq_hdr_len.assign(n_hdr, 0);
  if (n_hdr > pre_seqs) {
    const size_t cnt = (size_t)(n_hdr - pre_seqs);
    PM_CUDA(cudaMemcpyAsync(c->seq_key.data() + pre_seqs, c->d_fa_key.as<long long>() + pre_seqs, cnt * 8,
                            cudaMemcpyDeviceToHost, st));
    PM_CUDA(cudaMemcpyAsync(c->seq_hdr_off.data() + pre_seqs, c->d_fa_hs.as<long long>() + pre_seqs, cnt * 8,
                            cudaMemcpyDeviceToHost, st));
    PM_CUDA(cudaMemcpyAsync(c->seq_hdr_len.data() + pre_seqs, c->d_fa_he.as<long long>() + pre_seqs, cnt * 8,
                            cudaMemcpyDeviceToHost, st));
  }
  PM_CUDA(cudaStreamSynchronize(st));
  for (uint64_t i = 0; i < pre_seqs; i++) {
    c->seq_key[i] = pk[i];
    c->seq_hdr_off[i] = po[i];
    c->seq_hdr_len[i] = pl[i];
  }
  for (uint64_t i = pre_seqs; i < n_hdr; i++)
    c->seq_hdr_len[i] = c->seq_hdr_len[i] - c->seq_hdr_off[i] - 1;  // hend - hstart - 1 (fasta_io.t:548)
  c->d_text = d_stream;
  return set_text_common(c, n, 0);   // prep stream is drained: pack on the main stream
}

int pm_set_text_fasta_host(pm_ctx *c, const uint8_t *raw, uint64_t n_raw, int upper_case) {
  if (!c || (!raw && n_raw)) return fail(PM_E_ARG, "null argument");
  if (!c->inited) return fail(PM_E_ARG, "pm_set_text_fasta_host before pm_init");
  if (c->inner) return adopt_inner_text(c, pm_set_text_fasta_host(c->inner, raw, n_raw, upper_case));
  return set_text_fasta_impl(c, raw, n_raw, upper_case, nullptr);
}

static bool file_exists(const std::string &path) {
  struct stat st;
  return stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
static bool slurp_file(const std::string &path, std::string *out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  out->clear();
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}

// pick_fasta_file (select.t:23-190) + the reader: the database is opened the way primer_match
// opens it for this -D value and read through the ring of page-locked chunk buffers.
int pm_set_text_database(pm_ctx *c, const char *database, int dbind, int upper_case) {
  if (!c || !database) return fail(PM_E_ARG, "null argument");
  if (!c->inited) return fail(PM_E_ARG, "pm_set_text_database before pm_init");
  if (c->inner) return adopt_inner_text(c, pm_set_text_database(c->inner, database, dbind, upper_case));
  PM_CUDA(cudaSetDevice(c->opt.device));
  const std::string db(database);
  int kind = dbind;   // 1 raw FASTA, 2 .seq, 3 .sqn, 4 .sqz
  if (dbind == 0) kind = file_exists(db + ".sqn") ? 3 : (file_exists(db + ".sqz") ? 4 : (file_exists(db + ".seq") ? 2 : 1));
  if (kind < 1 || kind > 4) return fail(PM_E_ARG, "database selection (-D) must be 0..4");
  static const int n_threads = getenv("PM_READER_THREADS") ? std::max(1, atoi(getenv("PM_READER_THREADS"))) : 6;
  std::string err;
  c->db_kind = kind;
  delete c->reader;
  c->reader = new pm::ChunkReader();
  pm::ChunkReader &rd = *c->reader;
  if (kind == 1) {
    if (!rd.open(db, kFastaChunk, n_threads, n_threads + 2, &err)) return fail(PM_E_ARG, err);
    return set_text_fasta_impl(c, nullptr, rd.size(), upper_case, &rd);
  }
  // indexed databases of compress_seq: the stream (.seq), its table indices (.sqn + .tbl) or the
  // bit stream (.sqz + .tbz); sequence starts and header offsets in <db>.idb, headers in <db>.hdr
  std::string idb, table;
  if (!slurp_file(db + ".idb", &idb) || idb.size() < 8)
    return fail(PM_E_ARG, "cannot read " + db + ".idb (compress_seq's binary sequence index)");
  uint64_t n_ent = 0;
  memcpy(&n_ent, idb.data(), 8);
  if (n_ent < 1 || idb.size() < 8 + n_ent * 16) return fail(PM_E_ARG, db + ".idb is truncated");
  c->seq_key.clear();
  c->seq_hdr_off.clear();
  c->seq_hdr_len.clear();
  for (uint64_t i = 0; i + 1 < n_ent; i++) {   // the last entry is the end marker
    int64_t key = 0, hoff = 0, hnext = 0;
    memcpy(&key, idb.data() + 8 + i * 16, 8);
    memcpy(&hoff, idb.data() + 8 + i * 16 + 8, 8);
    memcpy(&hnext, idb.data() + 8 + (i + 1) * 16 + 8, 8);
    c->seq_key.push_back(key);
    c->seq_hdr_off.push_back(hoff);
    c->seq_hdr_len.push_back(hnext - hoff - 1);   // without the newline
  }
  if (!slurp_file(db + ".hdr", &c->hdr_blob)) return fail(PM_E_ARG, "cannot read " + db + ".hdr");
  c->hdr_ready = true;
  c->hdr_from_file = true;
  const std::string path = db + (kind == 2 ? ".seq" : (kind == 3 ? ".sqn" : ".sqz"));
  uint32_t bits = 0;
  if (kind != 2) {
    if (!slurp_file(db + (kind == 3 ? ".tbl" : ".tbz"), &table) || table.empty() || table.size() > 256)
      return fail(PM_E_ARG, "cannot read the character table of " + db);
    bits = (kind == 3) ? 8u : (uint32_t)pm_sqz_bits((uint32_t)table.size());
    uint8_t tab[256];
    memset(tab, 0, sizeof tab);
    memcpy(tab, table.data(), table.size());
    if (int rc = c->d_sqztab.ensure(256)) return rc;
    PM_CUDA(cudaMemcpyAsync(c->d_sqztab.p, tab, 256, cudaMemcpyHostToDevice, c->stream));
    PM_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (!rd.open(path, stream_chunk_bytes(bits), n_threads, n_threads + 2, &err)) return fail(PM_E_ARG, err);
  const uint64_t n_bytes = rd.size();
  const uint64_t n_chars = bits ? n_bytes * 8 / bits : n_bytes;   // the writer's padding reads as end-of-sequence characters
  if (n_chars >= (1ull << 40)) return fail(PM_E_UNSUPPORTED, "database longer than 2^40 characters");
  return stream_in(c, nullptr, n_bytes, bits, n_chars, 0, &rd);
}

// header line of sequence i (0-based; the text behind '>' without the line end): FastaFile's
// get_header_data (fasta_io.t:163-191).  For a FASTA database the header bytes are gathered from
// the raw file on the device once and kept; for an indexed database they come from <db>.hdr.
__global__ void gather_ranges_kernel(const uint8_t *src, const long long *off, const long long *len,
                                     const unsigned long long *out_off, uint64_t n, uint8_t *out) {
  const uint64_t i = blockIdx.x;
  if (i >= n) return;
  const long long l = len[i];
  for (long long x = threadIdx.x; x < l; x += blockDim.x) out[out_off[i] + x] = src[off[i] + x];
}

int pm_get_header(pm_ctx *c, uint64_t i, const char **text, uint64_t *len) {
  if (!c || !text || !len) return fail(PM_E_ARG, "null argument");
  pm_ctx *o = c->inner ? c->inner : c;   // whose pipeline brought the text in
  if (i >= o->seq_key.size()) return fail(PM_E_ARG, "sequence index out of range");
  if (!o->hdr_ready) {
    // FASTA text normalised on the device: the headers lie in the raw bytes there
    PM_CUDA(cudaSetDevice(o->opt.device));
    if (!o->d_fa_raw.p) return fail(PM_E_ARG, "headers are available for FASTA / indexed databases only");
    const size_t n = o->seq_key.size();
    std::vector<unsigned long long> ooff(n + 1, 0);
    std::vector<long long> hl(n);
    for (size_t x = 0; x < n; x++) {
      hl[x] = std::max<long long>(0, o->seq_hdr_len[x]);
      ooff[x + 1] = ooff[x] + (unsigned long long)hl[x];
    }
    LocalBuf d_off, d_len, d_ooff, d_out;
    int rc;
    if ((rc = d_off.ensure(n * 8 + 8)) || (rc = d_len.ensure(n * 8 + 8)) || (rc = d_ooff.ensure((n + 1) * 8)) ||
        (rc = d_out.ensure(ooff[n] + 8)))
      return rc;
    PM_CUDA(cudaMemcpyAsync(d_off.p, o->seq_hdr_off.data(), n * 8, cudaMemcpyHostToDevice, o->stream));
    PM_CUDA(cudaMemcpyAsync(d_len.p, hl.data(), n * 8, cudaMemcpyHostToDevice, o->stream));
    PM_CUDA(cudaMemcpyAsync(d_ooff.p, ooff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, o->stream));
    if (n) gather_ranges_kernel<<<(unsigned)n, 64, 0, o->stream>>>(o->d_fa_raw.as<uint8_t>(), d_off.as<long long>(),
                                                                   d_len.as<long long>(), d_ooff.as<unsigned long long>(),
                                                                   n, d_out.as<uint8_t>());
    o->hdr_blob.assign((size_t)ooff[n], '\0');
    if (ooff[n]) PM_CUDA(cudaMemcpyAsync(&o->hdr_blob[0], d_out.p, ooff[n], cudaMemcpyDeviceToHost, o->stream));
    PM_CUDA(cudaStreamSynchronize(o->stream));
    o->hdr_blob_off.assign(ooff.begin(), ooff.end());
    o->hdr_ready = true;
  }
  if (o->hdr_from_file) {
    const int64_t a = o->seq_hdr_off[i], l = std::max<int64_t>(0, o->seq_hdr_len[i]);
    if (a < 0 || (uint64_t)(a + l) > o->hdr_blob.size()) return fail(PM_E_ARG, "header index outside <db>.hdr");
    *text = o->hdr_blob.data() + a;
    *len = (uint64_t)l;
  } else {
    *text = o->hdr_blob.data() + o->hdr_blob_off[i];
    *len = o->hdr_blob_off[i + 1] - o->hdr_blob_off[i];
  }
  return PM_OK;
}

int pm_get_sequences(pm_ctx *c, const int64_t **seq_key, const int64_t **hdr_off, const int64_t **hdr_len,
                     uint64_t *n_seqs, uint64_t *stream_len) {
  if (!c || !n_seqs) return fail(PM_E_ARG, "null argument");
  if (seq_key) *seq_key = c->seq_key.data();
  if (hdr_off) *hdr_off = c->seq_hdr_off.data();
  if (hdr_len) *hdr_len = c->seq_hdr_len.data();
  *n_seqs = c->seq_key.size();
  if (stream_len) *stream_len = c->n;
  return PM_OK;
}

int pm_get_text(pm_ctx *c, uint64_t begin, uint64_t len, uint8_t *out) {
  if (!c || (!out && len)) return fail(PM_E_ARG, "null argument");
  if (begin + len > c->n || !c->d_text) return fail(PM_E_ARG, "range outside the loaded text");
  PM_CUDA(cudaSetDevice(c->opt.device));
  const pm_ctx *owner = c->inner ? c->inner : c;   // whose pipeline brought the text in
  if (!owner->chunks.empty() && owner->chunks.back().ev)
    PM_CUDA(cudaStreamWaitEvent(c->stream, owner->chunks.back().ev, 0));
  PM_CUDA(cudaMemcpyAsync(out, c->d_text + begin, len, cudaMemcpyDeviceToHost, c->stream));
  PM_CUDA(cudaStreamSynchronize(c->stream));
  return PM_OK;
}

// characters of n ranges of the loaded stream, concatenated: one gather on the device and one copy
// (the alignment buffers a tool prints next to its hits)
int pm_gather_text(pm_ctx *c, const int64_t *begin, const int64_t *len, uint64_t n, uint8_t *out) {
  if (!c || (n && (!begin || !len || !out))) return fail(PM_E_ARG, "null argument");
  if (!n) return PM_OK;
  if (!c->d_text) return fail(PM_E_ARG, "no text");
  PM_CUDA(cudaSetDevice(c->opt.device));
  std::vector<unsigned long long> ooff(n + 1, 0);
  for (uint64_t i = 0; i < n; i++) {
    if (begin[i] < 0 || len[i] < 0 || (uint64_t)(begin[i] + len[i]) > c->n) return fail(PM_E_ARG, "range outside the loaded text");
    ooff[i + 1] = ooff[i] + (unsigned long long)len[i];
  }
  const pm_ctx *owner = c->inner ? c->inner : c;
  if (!owner->chunks.empty() && owner->chunks.back().ev)
    PM_CUDA(cudaStreamWaitEvent(c->stream, owner->chunks.back().ev, 0));
  LocalBuf d_off, d_len, d_ooff, d_out;
  int rc;
  if ((rc = d_off.ensure(n * 8)) || (rc = d_len.ensure(n * 8)) || (rc = d_ooff.ensure((n + 1) * 8)) ||
      (rc = d_out.ensure(ooff[n] + 8)))
    return rc;
  PM_CUDA(cudaMemcpyAsync(d_off.p, begin, n * 8, cudaMemcpyHostToDevice, c->stream));
  PM_CUDA(cudaMemcpyAsync(d_len.p, len, n * 8, cudaMemcpyHostToDevice, c->stream));
  PM_CUDA(cudaMemcpyAsync(d_ooff.p, ooff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  gather_ranges_kernel<<<(unsigned)std::min<uint64_t>(n, 1u << 30), 32, 0, c->stream>>>(
      c->d_text, d_off.as<long long>(), d_len.as<long long>(), d_ooff.as<unsigned long long>(), n, d_out.as<uint8_t>());
  PM_CUDA(cudaGetLastError());
  if (ooff[n]) PM_CUDA(cudaMemcpyAsync(out, d_out.p, ooff[n], cudaMemcpyDeviceToHost, c->stream));
  PM_CUDA(cudaStreamSynchronize(c->stream));
  return PM_OK;
}

// Page-locked host memory for database buffers (what pm_set_text_*_host read directly).
// write_combined: the CPU only ever fills such a buffer front to back (read(2) / mmap copy),
// and the DMA engine then reads it without snooping the CPU caches.
int pm_host_alloc(uint64_t bytes, int write_combined, void **out) {
  if (!out) return fail(PM_E_ARG, "null argument");
  *out = nullptr;
  PM_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return PM_OK;
}

int pm_host_free(void *p) {
  if (p) PM_CUDA(cudaFreeHost(p));
  return PM_OK;
}

int pm_bench_int_peak(int device, double *lop3_tera_ops) {
  if (!lop3_tera_ops) return fail(PM_E_ARG, "null argument");
  PM_CUDA(cudaSetDevice(device));
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
  uint32_t *d = nullptr;
  PM_CUDA(cudaMalloc(&d, 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000, blocks = n_sm * 8;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, 0);
    pm::lop3_peak_kernel<<<blocks, 256>>>(d, iters);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256.0 * iters * 64.0;
    if (rep > 0 && ms > 0) best = std::max(best, ops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  PM_CUDA(cudaGetLastError());
  *lop3_tera_ops = best;
  return PM_OK;
}

int pm_pcr_pairs(const pm_hit *hits, uint64_t nhits, uint32_t n_primers, const uint32_t *patlen,
                 const int64_t *seq_key, uint64_t n_seqs, const pm_pcr_options *opt, pm_pcr_pair **pairs,
                 uint64_t *npairs) {
  if (!pairs || !npairs || !opt || !patlen || (nhits && !hits) || (n_seqs && !seq_key))
    return fail(PM_E_ARG, "null argument");
  *pairs = nullptr;
  *npairs = 0;
  if (n_primers == 0 || n_primers % 2 != 0) return fail(PM_E_ARG, "Odd number of primers!");  // pcr_match.cc:768
  std::vector<pm_pcr_pair> out;
  pcr_pairs(hits, nhits, n_primers, patlen, seq_key, n_seqs, *opt, out);
  if (!out.empty()) {
    *pairs = (pm_pcr_pair *)malloc(out.size() * sizeof(pm_pcr_pair));
    if (!*pairs) return fail(PM_E_NOMEM, "out of host memory");
    memcpy(*pairs, out.data(), out.size() * sizeof(pm_pcr_pair));
  }
  *npairs = out.size();
  return PM_OK;
}

void pm_free(void *p) { free(p); }

int pm_reverse_comp(const char *pattern, char *out) {
  if (!pattern || !out) return fail(PM_E_ARG, "null argument");
  const size_t n = strlen(pattern);
  for (size_t i = 0; i < n; i++) out[i] = (char)revcomp_char((unsigned char)pattern[n - 1 - i]);
  out[n] = 0;
  return PM_OK;
}

}  // extern "C"
