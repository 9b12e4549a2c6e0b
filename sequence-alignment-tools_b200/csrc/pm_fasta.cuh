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
// and state machines compose: every thread summarises 256 raw bytes as the map
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

__device__ __forceinline__ uint4 fa_load16(const uint8_t *raw, uint64_t at, uint64_t n) {
  if (at + 16 <= n && ((reinterpret_cast<uintptr_t>(raw + at) & 15) == 0))
    return __ldg(reinterpret_cast<const uint4 *>(raw + at));
  uint32_t w[4] = {0, 0, 0, 0};
  for (int j = 0; j < 16; j++)
    if (at + j < n) w[j >> 2] |= (uint32_t)raw[at + j] << (8 * (j & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// pass 1: per segment, the state map and the counts for each of the three start states
__global__ void __launch_bounds__(FA_THREADS)
fasta_summary_kernel(const uint8_t *__restrict__ raw, uint64_t r0, uint64_t n_raw, uint64_t n_seg,
                     uint8_t *__restrict__ fmap, unsigned long long *__restrict__ cnt_out,
                     uint32_t *__restrict__ cnt_hdr) {
  const uint64_t g = (uint64_t)blockIdx.x * FA_THREADS + threadIdx.x;
  if (g >= n_seg) return;
  const uint64_t begin = r0 + g * FA_SEG;
  uint32_t st[3] = {FA_SEQ, FA_HDR, FA_SKIP}, out[3] = {0, 0, 0}, hdr[3] = {0, 0, 0};
  for (int q = 0; q < FA_SEG / 16; q++) {
    const uint64_t at = begin + 16 * q;
    if (at >= n_raw) break;
    const uint4 v = fa_load16(raw, at, n_raw);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (at + j >= n_raw) break;
      const uint32_t b = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
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
__global__ void __launch_bounds__(256)
fasta_select_kernel(const uint8_t *__restrict__ fscan, const unsigned long long *__restrict__ cnt_out,
                    const uint32_t *__restrict__ cnt_hdr, uint64_t n_seg, uint8_t *__restrict__ start,
                    uint32_t *__restrict__ n_out, uint32_t *__restrict__ n_hdr) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_seg) return;
  const uint32_t s = g ? (fscan[g - 1] & 3u) : FA_SEQ;
  start[g] = (uint8_t)s;
  n_out[g] = (uint32_t)(cnt_out[g] >> (16 * s)) & 0xffffu;
  n_hdr[g] = (cnt_hdr[g] >> (8 * s)) & 0xffu;
}

// pass 2: write the stream and the sequence table.  out_base = stream index of the first
// byte the device writes; hdr_base = sequences registered by the host prefix.
__global__ void __launch_bounds__(FA_THREADS)
fasta_write_kernel(const uint8_t *__restrict__ raw, uint64_t r0, uint64_t n_raw, uint64_t n_seg,
                   const uint8_t *__restrict__ start, const unsigned long long *__restrict__ out_off,
                   const unsigned long long *__restrict__ hdr_off, int upper_case, uint32_t eos,
                   uint8_t *__restrict__ stream, uint64_t out_base, uint64_t hdr_base,
                   long long *__restrict__ seq_key, long long *__restrict__ hdr_start,
                   long long *__restrict__ hdr_end) {
  __shared__ uint8_t s_out[FA_THREADS * FA_SEG];
  __shared__ unsigned long long s_base, s_end;
  const uint64_t g0 = (uint64_t)blockIdx.x * FA_THREADS;
  const uint64_t g = g0 + threadIdx.x;
  if (threadIdx.x == 0) {
    s_base = out_off[g0];
    const uint64_t g_last = (g0 + FA_THREADS < n_seg) ? g0 + FA_THREADS : n_seg;
    s_end = out_off[g_last];  // out_off has n_seg + 1 entries
  }
  __syncthreads();
  if (g < n_seg) {
    const uint64_t begin = r0 + g * FA_SEG;
    uint32_t st = start[g];
    uint32_t o = (uint32_t)(out_off[g] - s_base);   // position inside the block's output
    uint64_t h = hdr_base + hdr_off[g];             // index of the next header to start
    for (int q = 0; q < FA_SEG / 16; q++) {
      const uint64_t at = begin + 16 * q;
      if (at >= n_raw) break;
      const uint4 v = fa_load16(raw, at, n_raw);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 16; j++) {
        if (at + j >= n_raw) break;
        uint32_t b = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
        if (st == FA_SEQ) {
          if (b == '>') {
            s_out[o] = (uint8_t)eos;
            seq_key[h] = (long long)(out_base + s_base + o + 1);  // first base of the sequence
            hdr_start[h] = (long long)(at + j + 1);
            h++;
            o++;
          } else if (!fa_blank(b)) {
            if (upper_case && b >= 'a' && b <= 'z') b -= 'a' - 'A';
            s_out[o++] = (uint8_t)b;
          }
        } else if (st == FA_HDR && (b == '\n' || b == '\r')) {
          hdr_end[h - 1] = (long long)(at + j + 1);   // the header in progress is the last one started
        }
        st = fa_next(st, b);
      }
    }
  }
  __syncthreads();
  const uint32_t total = (uint32_t)(s_end - s_base);
  uint8_t *dst = stream + out_base + s_base;
  for (uint32_t i = threadIdx.x; i < total; i += FA_THREADS) dst[i] = s_out[i];
}

__global__ void fill_i64_kernel(long long *p, uint64_t n, long long v) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace pm
