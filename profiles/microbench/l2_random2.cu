// Microbenchmark 2: random L2-resident probes with the sm_100a 256-bit load (LDG.256), L2 eviction hints, and a
// concurrent HBM stream (the real shape of the hash-probe kernel: stream 8 B in, probe, stream 4 B out).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_random2 l2_random2.cu && ./l2_random2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// MODE 0: 16B ld.ca   1: 32B ld (v4.b64)   2: 32B ld L2::evict_last   3: 16B ld.ca + L2 evict_last policy   4: 8B ld.ca
// STREAM: also read 8 B/row of a big array (nc, no_allocate) and write 4 B/row
template <int MODE, bool STREAM>
__global__ void __launch_bounds__(256) k(const void* table, uint32_t mask, int64_t n, const unsigned long long* in,
                                         uint32_t* out, unsigned long long* sink) {
  unsigned long long pol = 0;
  if (MODE == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  unsigned long long acc = 0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    unsigned long long kk[4] = {0, 0, 0, 0};
    if (STREAM) {
      const uint4* p = (const uint4*)(in + i);
      uint4 a, b;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 1));
      kk[0] = a.x; kk[1] = a.z; kk[2] = b.x; kk[3] = b.z;
    }
    uint32_t idx[4], o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) idx[j] = hash32((uint32_t)(i + j) ^ (uint32_t)kk[j]) & mask;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (MODE == 0) { uint4 v; asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"((const uint4*)table + idx[j])); o[j] = v.x ^ v.w; }
      if (MODE == 1) { unsigned long long a, b, c, d; asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"((const uint4*)table + 2 * (size_t)(idx[j] >> 1))); o[j] = (uint32_t)(a ^ d); }
      if (MODE == 2) { unsigned long long a, b, c, d; asm volatile("ld.global.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"((const uint4*)table + 2 * (size_t)(idx[j] >> 1))); o[j] = (uint32_t)(a ^ d); }
      if (MODE == 3) { uint4 v; asm volatile("ld.global.ca.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"((const uint4*)table + idx[j]), "l"(pol)); o[j] = v.x ^ v.w; }
      if (MODE == 4) { unsigned long long v; asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"((const unsigned long long*)table + idx[j])); o[j] = (uint32_t)v; }
    }
    if (STREAM) {
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + i), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
    } else {
      acc += o[0] ^ o[1] ^ o[2] ^ o[3];
    }
  }
  if (acc == 0x1234567) *sink = acc;
}

template <int MODE, bool STREAM>
void run(const char* name, void* table, unsigned long long* in, uint32_t* out, unsigned long long* sink) {
  printf("%-34s", name);
  for (int lg = 19; lg <= 23; lg += 1) {  // 16-byte elements: 8 MB .. 128 MB
    uint32_t mask = (1u << lg) - 1;
    int64_t n = 1ll << 29;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, STREAM><<<148 * 8, 256>>>(table, mask, n >> 3, in, out, sink);
    cudaEventRecord(a);
    k<MODE, STREAM><<<148 * 8, 256>>>(table, mask, n, in, out, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf(" %6.1f", n / (ms * 1e6));
  }
  printf("   G rows/s\n");
}

int main() {
  void* table; cudaMalloc(&table, (size_t)1 << 28); cudaMemset(table, 0, (size_t)1 << 28);
  unsigned long long* in; cudaMalloc(&in, (size_t)8 << 29); cudaMemset(in, 1, (size_t)8 << 29);
  uint32_t* out; cudaMalloc(&out, (size_t)4 << 29);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  printf("table bytes (16 B x 2^lg):          %6s %6s %6s %6s %6s\n", "8MB", "16MB", "32MB", "64MB", "128MB");
  run<0, false>("16B ld.ca", table, in, out, sink);
  run<3, false>("16B ld.ca L2 evict_last", table, in, out, sink);
  run<1, false>("32B ld.v4.b64 (LDG.256)", table, in, out, sink);
  run<2, false>("32B ld.v4.b64 L2 evict_last", table, in, out, sink);
  run<4, false>("8B ld.ca", table, in, out, sink);
  run<0, true>("16B ld.ca + stream", table, in, out, sink);
  run<3, true>("16B ld.ca evict_last + stream", table, in, out, sink);
  run<1, true>("32B LDG.256 + stream", table, in, out, sink);
  run<2, true>("32B LDG.256 evict_last + stream", table, in, out, sink);
  run<4, true>("8B ld.ca + stream", table, in, out, sink);
  return 0;
}
