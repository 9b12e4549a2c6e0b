// pm_planes.cuh -- text layout of the dense scan and the pack kernel (sm_100a).
//
// The reference feeds its indexes one character at a time through
// CharacterProducer::getnch() (char_io.h:18-71).  Here the normalised stream is
// turned ONCE into per-symbol match bit-planes, and every scan kernel works on
// those planes with all 32 bits of a word being 32 different text positions.
//
// Layout.  The stream is cut into tiles of TILE_POS = 32*TILE_W positions.
// Inside a tile, position  q = j*TILE_W + w  (0 <= w < TILE_W, 0 <= j < 32)
// lives in bit j of word w: a *transposed* layout.  The point of it: the
// character i places to the right of q is bit j of word w+i, so "shift the
// text by i" is an address offset, never a bit shift.  Words TILE_W ..
// TILE_WT-1 repeat the scheme one bit-row further (bit j of word TILE_W+e is
// position (j+1)*TILE_W+e; for j = 31 that is the first characters of the next
// tile), so a window of up to TILE_EXT+1 characters starting anywhere in the
// tile is addressable without wrap-around logic.
//
// Plane s (s < n_sym) has a 1 where the text character is matched by pattern
// symbol s under the scanner's relation (shift_and_inexact.cc:126-143,
// shift_and.cc:107-119); plane n_sym marks end-of-sequence characters and
// everything past the end of the stream.
//
// HBM layout: planes[tile][plane][TILE_WT] uint32, i.e. one tile's planes are
// contiguous (n_planes * 4352 bytes) so that a scan block fetches them with
// 1-D bulk copies (TMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

constexpr int TILE_W = 1024;
constexpr int TILE_EXT = 64;
constexpr int TILE_WT = TILE_W + TILE_EXT;
constexpr int TILE_POS = 32 * TILE_W;
constexpr int MAX_SYMBOLS = 30;  // symbol planes; +1 eos plane <= 31 mask bits
constexpr int PACK_THREADS = 288;  // >= TILE_WT/4 = 272 threads, 4 words each
constexpr int PACK_GROUP = 8;      // planes accumulated per pass over the tile

// symmask[b]: bit s set <=> text byte b is matched by pattern symbol s;
// bit n_sym set <=> b is the end-of-sequence character.
__global__ void __launch_bounds__(PACK_THREADS)
pack_planes_kernel(const uint8_t *__restrict__ text, uint64_t n,
                   const uint32_t *__restrict__ symmask, int n_planes, uint32_t eos_byte,
                   uint32_t *__restrict__ planes, uint32_t tile_base) {
  __shared__ uint32_t s_mask[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_mask[i] = symmask[i];
  __syncthreads();
  const uint32_t q = threadIdx.x;
  if (q >= TILE_WT / 4) return;
  const uint64_t tile = (uint64_t)tile_base + blockIdx.x;
  const uint64_t base = tile * (uint64_t)TILE_POS + 4u * q;
  const uint32_t eos4 = eos_byte * 0x01010101u;
  for (int p0 = 0; p0 < n_planes; p0 += PACK_GROUP) {
    uint32_t acc[PACK_GROUP][4];
#pragma unroll
    for (int s = 0; s < PACK_GROUP; s++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[s][b] = 0;
#pragma unroll 4
    for (int j = 0; j < 32; j++) {
      const uint64_t idx = base + (uint64_t)j * TILE_W;
      uint32_t v;
      if (idx + 4 <= n) {
        v = __ldg(reinterpret_cast<const uint32_t *>(text + idx));
      } else {
        v = eos4;  // past the end of the stream reads as end-of-sequence
        for (int b = 0; b < 4; b++)
          if (idx + b < n) v = (v & ~(0xffu << (8 * b))) | ((uint32_t)text[idx + b] << (8 * b));
      }
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t m = s_mask[(v >> (8 * b)) & 0xffu] >> p0;
#pragma unroll
        for (int s = 0; s < PACK_GROUP; s++) acc[s][b] |= ((m >> s) & 1u) << j;
      }
    }
#pragma unroll
    for (int s = 0; s < PACK_GROUP; s++) {
      if (p0 + s < n_planes) {
        uint4 *dst = reinterpret_cast<uint4 *>(planes + (tile * n_planes + (p0 + s)) * (uint64_t)TILE_WT) + q;
        *dst = make_uint4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Compressed<> (char_io.t:18-215): the .sqz bit stream compress_seq -z writes
// (compress_seq.cc:745-870) -> stream bytes.  `bits` per character, most
// significant bit first, character i at bit i*bits of the byte stream; table =
// contents of the .tbz file.  One thread turns 2*bits bytes into 16 characters
// (16 characters always start on a byte boundary) and stores them as one uint4.
// Characters at or past n_chars are written as end-of-sequence (the writer pads
// its last buffer with them anyway).
constexpr int SQZ_GROUPS_PER_THREAD = 4;
template <int BITS>
__global__ void __launch_bounds__(256)
unpack_sqz_kernel(const uint8_t *__restrict__ sqz, uint64_t n_bytes, const uint8_t *__restrict__ table,
                  uint32_t eos_byte, uint8_t *__restrict__ text, uint64_t g_begin, uint64_t g_end,
                  uint64_t n_chars) {
  constexpr uint32_t bits = BITS;
  __shared__ uint8_t s_tab[256];
  s_tab[threadIdx.x] = table[threadIdx.x];
  __syncthreads();
  const uint32_t mask = (1u << bits) - 1u;
  const uint64_t g_first = g_begin + (uint64_t)blockIdx.x * (256 * SQZ_GROUPS_PER_THREAD) + threadIdx.x;
#pragma unroll 1
  for (int r = 0; r < SQZ_GROUPS_PER_THREAD; r++) {
    const uint64_t g = g_first + (uint64_t)r * 256;  // group of 16 characters
    if (g >= g_end) return;
    const uint64_t b0 = g * 2ull * bits;
    // 2*bits (<= 16) payload bytes, big-endian, plus one byte of slack for the 16-bit window
    uint32_t by[5] = {0, 0, 0, 0, 0};  // by[j] = bytes 4j .. 4j+3, most significant first
#pragma unroll
    for (int j = 0; j < 2 * BITS; j++) {
      const uint64_t at = b0 + j;
      const uint32_t v = (at < n_bytes) ? (uint32_t)__ldg(sqz + at) : 0u;
      by[j >> 2] |= v << (24 - 8 * (j & 3));
    }
    uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t bit = (uint32_t)i * bits;        // < 128, known at compile time
      const uint32_t wj = bit >> 5, sh = bit & 31u;
      // 32-bit window starting at `bit` (big-endian bit order)
      const uint32_t hi = by[wj], lo = by[wj + 1 < 5 ? wj + 1 : 4];
      const uint32_t win = sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
      const uint32_t code = (win >> (32u - bits)) & mask;
      uint32_t ch = s_tab[code];
      if (g * 16 + i >= n_chars) ch = eos_byte;
      out[i >> 2] |= ch << (8 * (i & 3));
    }
    *reinterpret_cast<uint4 *>(text + g * 16) = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

}  // namespace pm
