// Microbenchmark: ceilings for random L2-resident access on B200 (what bounds the hash probe and the atomic reduce).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_random l2_random.cu && ./l2_random
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>  // 0: 16B ldcg, 1: 32B (2x16B) ldcg, 2: f64 atomicAdd, 3: u64 atomicMin, 4: u32 atomicAdd, 5: 16B ld.ca(default), 6: 8B ldcg, 7: 4B ldcg
__global__ void __launch_bounds__(256) k(void* table, uint32_t mask, int64_t n, unsigned long long* sink) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * 4) {
    uint32_t idx[4];
#pragma unroll
    for (int j = 0; j < 4; j++) idx[j] = hash32((uint32_t)(i + j * stride)) & mask;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (MODE == 0) { uint4 v = __ldcg((const uint4*)table + idx[j]); acc += v.x ^ v.w; }
      if (MODE == 1) { uint4 v = __ldcg((const uint4*)table + 2 * (idx[j] >> 1)); uint4 w = __ldcg((const uint4*)table + 2 * (idx[j] >> 1) + 1); acc += v.x ^ w.w; }
      if (MODE == 2) atomicAdd((double*)table + idx[j], 1.0);
      if (MODE == 3) atomicMin((unsigned long long*)table + idx[j], (unsigned long long)i);
      if (MODE == 4) atomicAdd((uint32_t*)table + idx[j], 1u);
      if (MODE == 5) { uint4 v = *((const uint4*)table + idx[j]); acc += v.x ^ v.w; }
      if (MODE == 6) { acc += __ldcg((const unsigned long long*)table + idx[j]); }
      if (MODE == 7) { acc += __ldcg((const uint32_t*)table + idx[j]); }
      if (MODE == 8) { uint4 v = __ldg((const uint4*)table + idx[j]); acc += v.x ^ v.w; }
      if (MODE == 9) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"((const uint4*)table + idx[j])); acc += v.x ^ v.w; }
      if (MODE == 10) { acc += *((const unsigned long long*)table + idx[j]); }
      if (MODE == 11) { ((uint32_t*)table)[idx[j]] = (uint32_t)i; }
      if (MODE == 12) { uint4 v = *((const uint4*)table + idx[j]); acc += v.x ^ v.w; if ((v.y & 0xff) == 0x7f) atomicAdd((double*)table + (idx[j] ^ 1) * 2, 1.0); }
      if (MODE == 13) { const unsigned long long* p = (const unsigned long long*)table + 2 * (size_t)idx[j]; acc += p[0] ^ p[1]; }
    }
  }
  if (acc == 0x1234567) *sink = acc;
}

template <int MODE>
void run(const char* name, void* table, size_t elem_bytes, unsigned long long* sink) {
  printf("%-14s", name);
  for (int lg = 16; lg <= 26; lg += 1) {  // elements
    uint32_t mask = (1u << lg) - 1;
    int64_t n = 1ll << 31;
    if ((MODE >= 2 && MODE <= 4) || MODE == 11) n = 1ll << 30;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 8, 256>>>(table, mask, n >> 3, sink);
    cudaEventRecord(a);
    k<MODE><<<148 * 8, 256>>>(table, mask, n, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf(" %6.1f", n / (ms * 1e6));
    (void)elem_bytes;
  }
  printf("   Gops/s\n");
}

int main() {
  void* table; cudaMalloc(&table, (size_t)1 << 31); cudaMemset(table, 0, (size_t)1 << 31);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  printf("elements(log2):"); for (int lg = 16; lg <= 26; lg++) printf(" %6d", lg); printf("\n");
  run<0>("ld16 cg", table, 16, sink);
  run<5>("ld16 ca", table, 16, sink);
  run<1>("ld32 cg", table, 16, sink);
  run<6>("ld8 cg", table, 8, sink);
  run<7>("ld4 cg", table, 4, sink);
  run<2>("atomAdd f64", table, 8, sink);
  run<3>("atomMin u64", table, 8, sink);
  run<4>("atomAdd u32", table, 4, sink);
  run<8>("ld16 ldg(nc)", table, 16, sink);
  run<9>("ld16 nc.noalloc", table, 16, sink);
  run<10>("ld8 ca", table, 8, sink);
  run<13>("2x ld8 ca", table, 16, sink);
  run<11>("st4 random", table, 4, sink);
  return 0;
}
