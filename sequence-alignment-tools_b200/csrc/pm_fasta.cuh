// pm_fasta.cuh -- FASTA normalisation on the device (sm_100a).
//
// StreamedFastaFile::getch (fasta_io.t:478-629, eos_start = true) turns a FASTA
// file into the stream  eos, seq1, eos, seq2, ..., seqN, eos : sequence lines are
// joined, blanks / CR / LF inside sequence data are dropped, '>' starts a header
// that runs to the end of the line.  The reference does this one character at a
// time; as a function of the raw bytes it is a three-state machine
//
//     SEQ   '>'  -> HDR (emits the separating eos)   blank -> dropped   else -> kept
//     HDR   '\n' -> SEQ    '\r' -> SKIP              else  -> header text
//     SKIP  any  -> SEQ  (the byte after a CR that ended a header is swallowed,
//                         fasta_io.t:532-533, whatever it is)
//
// and state machines compose: the file is taken in chunks as they arrive over PCIe (the
// state, the stream length and the header count after a chunk are carried in device
// memory); within a chunk every thread summarises 256 raw bytes as the map
// "state before -> state after" plus, per start state, how many stream bytes and
// headers it produces (fasta_summary_kernel); an inclusive scan of the maps under
// composition (CUB, 1 byte per segment) gives every segment its start state; two
// exclusive sums give it its place in the stream and in the sequence table; a
// second pass over the raw bytes writes both (fasta_write_kernel, staged through
// shared memory so that the stream is written in coalesced runs).
//
// The one non-local rule of the reference -- the character after the FIRST header
// line is taken verbatim, and no eos is emitted for that header (fasta_io.t:535-560)
// -- concerns a few bytes at the start of the file: the host walks the raw bytes
// until the first stream character has been produced and the device starts behind
// that point in state SEQ.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

constexpr int FA_SEG = 256;      // raw bytes per thread
constexpr int FA_THREADS = 128;  // 32 KB of raw bytes per block
enum : uint32_t { FA_SEQ = 0, FA_HDR = 1, FA_SKIP = 2 };

__device__ __forceinline__ uint32_t fa_next(uint32_t s, uint32_t b) {
  if (s == FA_SEQ) return b == '>' ? FA_HDR : FA_SEQ;
  if (s == FA_HDR) return b == '\n' ? FA_SEQ : (b == '\r' ? FA_SKIP : FA_HDR);
  return FA_SEQ;
}
__device__ __forceinline__ bool fa_blank(uint32_t b) { return b == '\n' || b == '\r' || b == ' '; }

// state maps are 6 bits: next state from SEQ in bits 0-1, from HDR in 2-3, from SKIP in 4-5
struct FaCompose {
  __host__ __device__ __forceinline__ uint8_t operator()(uint8_t first, uint8_t then) const {
    const uint32_t a = first, b = then;
    const uint32_t r0 = (b >> (2 * (a & 3u))) & 3u;
    const uint32_t r1 = (b >> (2 * ((a >> 2) & 3u))) & 3u;
    const uint32_t r2 = (b >> (2 * ((a >> 4) & 3u))) & 3u;
    return (uint8_t)(r0 | (r1 << 2) | (r2 << 4));
  }
};

// A block owns FA_THREADS * FA_SEG = 32 KB of raw bytes.  They are brought into shared
// memory with coalesced 16-byte loads; thread t then walks its own 256-byte segment, which
// sits at byte t * (FA_SEG + 4) so that the 32 lanes of a warp read 32 different banks.
constexpr int FA_ROW = FA_SEG + 4;   // padded segment row
constexpr int FA_STAGE_BYTES = FA_THREADS * FA_ROW;

__device__ __forceinline__ void fa_stage_block(const uint8_t *__restrict__ raw, uint64_t begin, uint64_t n_raw,
                                               uint32_t *s_in) {
  // begin = first raw byte of the block; bytes at or past n_raw read as 0 and are never used
  const uint32_t mis = (uint32_t)((uintptr_t)(raw + begin) & 15u);
  if (mis == 0 && begin + (uint64_t)FA_THREADS * FA_SEG <= n_raw) {
    const uint4 *src = reinterpret_cast<const uint4 *>(raw + begin);
    for (uint32_t i = threadIdx.x; i < FA_THREADS * FA_SEG / 16; i += FA_THREADS) {
      const uint4 v = __ldg(src + i);
      const uint32_t seg = i / (FA_SEG / 16), q = i % (FA_SEG / 16);
      uint32_t *dst = s_in + seg * (FA_ROW / 4) + q * 4;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
  } else {  // unaligned start (the host prefix ends anywhere) or the last block: word by word
    for (uint32_t i = threadIdx.x; i < FA_THREADS * FA_SEG / 4; i += FA_THREADS) {
      const uint64_t at = begin + 4ull * i;
      uint32_t v = 0;
      for (int j = 0; j < 4; j++)
        if (at + j < n_raw) v |= (uint32_t)raw[at + j] << (8 * j);
      s_in[(i / (FA_SEG / 4)) * (FA_ROW / 4) + (i % (FA_SEG / 4))] = v;
    }
  }
  __syncthreads();
}

// pass 1: per segment, the state map and the counts for each of the three start states
__global__ void __launch_bounds__(FA_THREADS)
fasta_summary_kernel(const uint8_t *__restrict__ raw, uint64_t r0, uint64_t n_raw, uint64_t n_seg,
                     uint8_t *__restrict__ fmap, unsigned long long *__restrict__ cnt_out,
                     uint32_t *__restrict__ cnt_hdr) {
  __shared__ uint32_t s_in[FA_STAGE_BYTES / 4];
  const uint64_t g0 = (uint64_t)blockIdx.x * FA_THREADS;
  fa_stage_block(raw, r0 + g0 * FA_SEG, n_raw, s_in);
  const uint64_t g = g0 + threadIdx.x;
  if (g >= n_seg) return;
  const uint64_t begin = r0 + g * FA_SEG;
  const int len = (int)((n_raw - begin < (uint64_t)FA_SEG) ? (n_raw - begin) : (uint64_t)FA_SEG);
  const uint32_t *row = s_in + threadIdx.x * (FA_ROW / 4);
  uint32_t st[3] = {FA_SEQ, FA_HDR, FA_SKIP}, out[3] = {0, 0, 0}, hdr[3] = {0, 0, 0};
  for (int q = 0; q < FA_SEG / 4; q++) {
    if (4 * q >= len) break;
    const uint32_t w = row[q];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (4 * q + j >= len) break;
      const uint32_t b = (w >> (8 * j)) & 0xffu;
      const bool gt = (b == '>'), blank = fa_blank(b);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (st[k] == FA_SEQ) {
          out[k] += blank ? 0u : 1u;   // a kept character, or the eos a header starts with
          hdr[k] += gt ? 1u : 0u;
        }
        st[k] = fa_next(st[k], b);
      }
    }
  }
  fmap[g] = (uint8_t)(st[0] | (st[1] << 2) | (st[2] << 4));
  cnt_out[g] = (unsigned long long)out[0] | ((unsigned long long)out[1] << 16) | ((unsigned long long)out[2] << 32);
  cnt_hdr[g] = hdr[0] | (hdr[1] << 8) | (hdr[2] << 16);
}

// between the scans: the counts that belong to the segment's actual start state
// The file is processed in chunks as it arrives; what one chunk hands to the next lives in
// device memory, so no chunk waits for the host: carry[0] = state after the last byte,
// carry[1] = stream bytes written, carry[2] = headers started.
__global__ void __launch_bounds__(256)
fasta_select_kernel(const uint8_t *__restrict__ fscan, const unsigned long long *__restrict__ cnt_out,
                    const uint32_t *__restrict__ cnt_hdr, uint64_t n_seg, uint8_t *__restrict__ start,
                    uint32_t *__restrict__ n_out, uint32_t *__restrict__ n_hdr,
                    const unsigned long long *__restrict__ carry) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_seg) return;
  const uint32_t s0 = (uint32_t)carry[0];                    // state the chunk starts in
  const uint32_t s = g ? ((fscan[g - 1] >> (2 * s0)) & 3u) : s0;
  start[g] = (uint8_t)s;
  n_out[g] = (uint32_t)(cnt_out[g] >> (16 * s)) & 0xffffu;
  n_hdr[g] = (cnt_hdr[g] >> (8 * s)) & 0xffu;
}

// pass 2: write the stream and the sequence table.  out_base = stream index of the first
// byte the device writes; hdr_base = sequences registered by the host prefix.  The block's
// output is collected in shared memory at the same 16-byte phase as its place in the
// stream, so that everything but its first and last few bytes leaves as aligned uint4.
constexpr int FA_WRITE_SMEM = FA_STAGE_BYTES + FA_THREADS * FA_SEG + 32;
__global__ void __launch_bounds__(FA_THREADS)
fasta_write_kernel(const uint8_t *__restrict__ raw, uint64_t r0, uint64_t n_raw, uint64_t n_seg,
                   const uint8_t *__restrict__ start, const unsigned long long *__restrict__ out_off,
                   const unsigned long long *__restrict__ hdr_off, int upper_case, uint32_t eos,
                   uint8_t *__restrict__ stream, uint64_t out_base0, uint64_t hdr_base0,
                   long long *__restrict__ seq_key, long long *__restrict__ hdr_start,
                   long long *__restrict__ hdr_end, const unsigned long long *__restrict__ carry,
                   uint64_t hdr_cap, unsigned long long *__restrict__ hdr_overflow) {
  const uint64_t out_base = out_base0 + carry[1];   // + what earlier chunks wrote
  const uint64_t hdr_base = hdr_base0 + carry[2];
  extern __shared__ __align__(16) uint32_t s_dyn[];
  uint32_t *s_in = s_dyn;
  uint8_t *s_out = reinterpret_cast<uint8_t *>(s_dyn + FA_STAGE_BYTES / 4);
  __shared__ unsigned long long s_base, s_end;
  const uint64_t g0 = (uint64_t)blockIdx.x * FA_THREADS;
  const uint64_t g = g0 + threadIdx.x;
  if (threadIdx.x == 0) {
    s_base = out_off[g0];
    const uint64_t g_last = (g0 + FA_THREADS < n_seg) ? g0 + FA_THREADS : n_seg;
    s_end = out_off[g_last];  // out_off has n_seg + 1 entries
  }
  fa_stage_block(raw, r0 + g0 * FA_SEG, n_raw, s_in);   // ends with __syncthreads()
  uint8_t *dst = stream + out_base + s_base;             // where the block's output goes
  const uint32_t phase = (uint32_t)((uintptr_t)dst & 15u);
  if (g < n_seg) {
    const uint64_t begin = r0 + g * FA_SEG;
    const int len = (int)((n_raw - begin < (uint64_t)FA_SEG) ? (n_raw - begin) : (uint64_t)FA_SEG);
    const uint32_t *row = s_in + threadIdx.x * (FA_ROW / 4);
    uint32_t st = start[g];
    uint32_t o = phase + (uint32_t)(out_off[g] - s_base);   // position inside s_out
    uint64_t h = hdr_base + hdr_off[g];                     // index of the next header to start
    for (int q = 0; q < FA_SEG / 4; q++) {
      if (4 * q >= len) break;
      const uint32_t w = row[q];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (4 * q + j >= len) break;
        uint32_t b = (w >> (8 * j)) & 0xffu;
        if (st == FA_SEQ) {
          if (b == '>') {
            s_out[o] = (uint8_t)eos;
            if (h < hdr_cap) {
              seq_key[h] = (long long)(out_base + s_base + (o - phase) + 1);  // first base of the sequence
              hdr_start[h] = (long long)(begin + 4 * q + j + 1);
            } else {
              *hdr_overflow = 1ULL;
            }
            h++;
            o++;
          } else if (!fa_blank(b)) {
            if (upper_case && b >= 'a' && b <= 'z') b -= 'a' - 'A';
            s_out[o++] = (uint8_t)b;
          }
        } else if (st == FA_HDR && (b == '\n' || b == '\r')) {
          if (h - 1 < hdr_cap) hdr_end[h - 1] = (long long)(begin + 4 * q + j + 1);   // the header in progress
        }
        st = fa_next(st, b);
      }
    }
  }
  __syncthreads();
  // s_out[phase .. phase+total) -> dst[0 .. total); s_out index i and address dst - phase + i
  // have the same 16-byte phase
  const uint32_t total = (uint32_t)(s_end - s_base);
  const uint32_t lo = phase, hi = phase + total;
  const uint32_t a16 = (lo + 15u) & ~15u, b16 = hi & ~15u;   // aligned interior [a16, b16)
  uint8_t *base = dst - phase;
  if (a16 < b16) {
    for (uint32_t i = lo + threadIdx.x; i < a16; i += FA_THREADS) base[i] = s_out[i];
    for (uint32_t i = a16 / 16 + threadIdx.x; i < b16 / 16; i += FA_THREADS)
      reinterpret_cast<uint4 *>(base)[i] = reinterpret_cast<const uint4 *>(s_out)[i];
    for (uint32_t i = b16 + threadIdx.x; i < hi; i += FA_THREADS) base[i] = s_out[i];
  } else {
    for (uint32_t i = lo + threadIdx.x; i < hi; i += FA_THREADS) base[i] = s_out[i];
  }
}

// after a chunk: fold its totals into the carry
__global__ void fasta_carry_kernel(const uint8_t *__restrict__ fscan, uint64_t n_seg,
                                   const unsigned long long *__restrict__ out_off,
                                   const unsigned long long *__restrict__ hdr_off, unsigned long long *carry) {
  if (threadIdx.x || blockIdx.x) return;
  const uint32_t s0 = (uint32_t)carry[0];
  carry[0] = (fscan[n_seg - 1] >> (2 * s0)) & 3u;
  carry[1] += out_off[n_seg];
  carry[2] += hdr_off[n_seg];
}

__global__ void fill_i64_kernel(long long *p, uint64_t n, long long v) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace pm
