// ubench_pipes.cu -- issue-rate microbenchmarks for the integer instructions the
// pair-seed scan is built from (sm_100a).  Prints thread-ops per clock per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t *out, int iters, uint32_t m, uint32_t s) {
  uint32_t a[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; j++) a[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int j = 0; j < CHAINS; j++) {
        uint32_t x = a[j], y;
        if (OP == 0) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(y) : "r"(x), "r"(m), "r"(s));
        if (OP == 1) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(x), "r"(m), "r"(s));
        if (OP == 2) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(y) : "r"(x), "r"(m));
        if (OP == 3) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(m)); y = (uint32_t)(w >> 32) ^ (uint32_t)w; }
        if (OP == 4) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(y) : "r"(x), "r"(m), "r"(s));
        if (OP == 5) {  // LOP3 + IMAD pair
          uint32_t t;
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(t) : "r"(x), "r"(m), "r"(s));
          asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(t), "r"(m), "r"(s));
        }
        if (OP == 6) {  // SHF + mul.hi pair
          uint32_t t;
          asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(x), "r"(m), "r"(s));
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(y) : "r"(t), "r"(m));
        }
        if (OP == 7) {  // 2 ALU + 1 IMAD + 1 mul.hi
          uint32_t t, u, v;
          asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(x), "r"(m), "r"(s));
          asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(t), "r"(m), "r"(s));
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(v) : "r"(u), "r"(m));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(y) : "r"(v), "r"(m), "r"(x));
        }
        if (OP == 8) asm volatile("mad.lo.u32 %0, %1, 4, %2;" : "=r"(y) : "r"(x), "r"(s));   // address form
        if (OP == 9) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(y) : "r"(x), "r"(m), "r"(s));
        a[j] = y;
      }
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; j++) x ^= a[j];
  if (x == 0x1234567u) out[0] = x;
}

template <int OP>
void run(const char *name, int ops_per_iter, int n_sm, double mhz) {
  uint32_t *d;
  cudaMalloc(&d, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4000;
  k<OP><<<n_sm, 1024>>>(d, 100, 0x9E3779B1u, 7);
  cudaEventRecord(e0);
  k<OP><<<n_sm, 1024>>>(d, iters, 0x9E3779B1u, 7);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)n_sm * 1024.0 * iters * 4.0 * CHAINS * ops_per_iter;
  printf("%-28s %8.3f ms  %7.2f thread-ops/clk/SM (at %.0f MHz)  %6.2f Tops/s\n", name, ms,
         ops / (ms * 1e-3) / (mhz * 1e6) / n_sm, mhz, ops / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

int main() {
  int n_sm = 148, khz = 1965000;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  run<0>("lop3", 1, n_sm, mhz);
  run<1>("mad.lo (IMAD)", 1, n_sm, mhz);
  run<2>("mul.hi (IMAD.HI)", 1, n_sm, mhz);
  run<9>("mad.hi (IMAD.HI 3-op)", 1, n_sm, mhz);
  run<3>("mul.wide (IMAD.WIDE)", 1, n_sm, mhz);
  run<4>("shf.r.wrap", 1, n_sm, mhz);
  run<8>("mad.lo x4+base", 1, n_sm, mhz);
  run<5>("lop3+mad.lo", 2, n_sm, mhz);
  run<6>("shf+mul.hi", 2, n_sm, mhz);
  run<7>("shf+mad+mul.hi+lop3", 4, n_sm, mhz);
  return 0;
}
