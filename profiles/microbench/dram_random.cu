// Microbenchmark 5: random accesses to tables far larger than L2 (the high-cardinality grouping regime), and the
// building blocks of a hash-partitioned grouping: a one-pass partition through global cursors, random 4-byte stores to row
// order, random bit sets.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dram_random dram_random.cu && ./dram_random
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}

__device__ __forceinline__ void load4keys(const unsigned long long* in, int64_t i, unsigned long long kk[4]) {
  const ulonglong2* p = (const ulonglong2*)(in + i);
  ulonglong2 a = __ldg(p), b = __ldg(p + 1);
  kk[0] = a.x; kk[1] = a.y; kk[2] = b.x; kk[3] = b.y;
}

// MODE 0: 32 B read per row (LDG.256) + 8 B in + 4 B out    1: 4 B store per row + 8 B in    2: 8 B store per row + 8 B in
// 3: 4 B read per row + 8 B in + 4 B out
template <int MODE>
__global__ void __launch_bounds__(256) rnd(void* table, uint64_t mask, int64_t n, const unsigned long long* in, uint32_t* out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    unsigned long long kk[4];
    load4keys(in, i, kk);
    uint64_t idx[4];
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) idx[j] = mix64((uint64_t)(i + j) + kk[j]) & mask;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (MODE == 0) {
        unsigned long long a, b, c, d;
        asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"((const char*)table + idx[j] * 32));
        o[j] = (uint32_t)(a ^ d);
      }
      if (MODE == 1) ((uint32_t*)table)[idx[j]] = (uint32_t)(i + j);
      if (MODE == 2) ((unsigned long long*)table)[idx[j]] = (unsigned long long)(i + j);
      if (MODE == 3) o[j] = __ldg((const uint32_t*)table + idx[j]);
    }
    if (MODE == 0 || MODE == 3) *(uint4*)(out + i) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one-pass partition through global cursors: pos = base(p) + atomicAdd(cursor[p], 1)
__global__ void __launch_bounds__(256) part(const unsigned long long* in, int64_t n, int lgp, uint32_t cap, uint32_t* cursor,
                                            unsigned long long* keys_out, uint32_t* rows_out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    unsigned long long kk[4];
    load4keys(in, i, kk);
    uint32_t p[4], r[4];
#pragma unroll
    for (int j = 0; j < 4; j++) p[j] = (uint32_t)(mix64(kk[j] + (uint64_t)(i + j)) >> (64 - lgp));
#pragma unroll
    for (int j = 0; j < 4; j++) r[j] = atomicAdd(cursor + p[j], 1u);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (r[j] < cap) {
        size_t pos = (size_t)p[j] * cap + r[j];
        keys_out[pos] = kk[j];
        rows_out[pos] = (uint32_t)(i + j);
      }
    }
  }
}

__global__ void __launch_bounds__(256) bits(uint32_t* bitmap, uint64_t mask, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t b = mix64(i) & mask;
    atomicOr(bitmap + (b >> 5), 1u << (b & 31));
  }
}

static float timeit(void (*launch)(void*), void* ctx) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(ctx);
  cudaEventRecord(a);
  launch(ctx);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  return ms;
}

struct Ctx { int mode; void* table; uint64_t mask; int64_t n; unsigned long long* in; uint32_t* out; int lgp; uint32_t cap; uint32_t* cursor; unsigned long long* ko; uint32_t* ro; };

static void launch_rnd(void* c_) {
  Ctx* c = (Ctx*)c_;
  if (c->mode == 0) rnd<0><<<148 * 8, 256>>>(c->table, c->mask, c->n, c->in, c->out);
  if (c->mode == 1) rnd<1><<<148 * 8, 256>>>(c->table, c->mask, c->n, c->in, c->out);
  if (c->mode == 2) rnd<2><<<148 * 8, 256>>>(c->table, c->mask, c->n, c->in, c->out);
  if (c->mode == 3) rnd<3><<<148 * 8, 256>>>(c->table, c->mask, c->n, c->in, c->out);
}
static void launch_part(void* c_) {
  Ctx* c = (Ctx*)c_;
  cudaMemsetAsync(c->cursor, 0, sizeof(uint32_t) << c->lgp);
  part<<<148 * 8, 256>>>(c->in, c->n, c->lgp, c->cap, c->cursor, c->ko, c->ro);
}
static void launch_bits(void* c_) {
  Ctx* c = (Ctx*)c_;
  bits<<<148 * 8, 256>>>((uint32_t*)c->table, c->mask, c->n);
}

int main(int argc, char** argv) {
  // optional argument: the L2 fetch granularity hint in bytes (cudaLimitMaxL2FetchGranularity: 32, 64 or 128)
  if (argc > 1) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity: asked %s, set -> %s, now %zu\n", argv[1], cudaGetErrorString(e), got);
  } else {
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity (default): %zu\n", got);
  }
  const int64_t n = 1ll << 29;
  void* table; cudaMalloc(&table, (size_t)16 << 30); cudaMemset(table, 0, (size_t)16 << 30);
  unsigned long long* in; cudaMalloc(&in, (size_t)8 * n); cudaMemset(in, 1, (size_t)8 * n);
  uint32_t* out; cudaMalloc(&out, (size_t)4 * n);
  Ctx c{}; c.table = table; c.n = n; c.in = in; c.out = out;
  const char* names[4] = {"32 B read (LDG.256) + 8 B in + 4 B out", "4 B store + 8 B in", "8 B store + 8 B in", "4 B read + 8 B in + 4 B out"};
  const int unit[4] = {32, 4, 8, 4};
  printf("random access, 2^29 rows, G rows/s by table size\n%-42s %8s %8s %8s %8s %8s\n", "", "256MB", "1GB", "4GB", "8GB", "16GB");
  for (int mode = 0; mode < 4; mode++) {
    printf("%-42s", names[mode]);
    const size_t sizes[5] = {(size_t)256 << 20, (size_t)1 << 30, (size_t)4 << 30, (size_t)8 << 30, (size_t)16 << 30};
    for (int s = 0; s < 5; s++) {
      c.mode = mode; c.mask = sizes[s] / unit[mode] - 1;
      float ms = timeit(launch_rnd, &c);
      printf(" %8.1f", n / (ms * 1e6));
    }
    printf("\n");
  }
  printf("one-pass partition of (8 B key, 4 B row) through global cursors, 2^29 rows\n");
  unsigned long long* ko = (unsigned long long*)table;
  for (int lgp = 8; lgp <= 16; lgp += 2) {
    uint32_t cap = (uint32_t)((n >> lgp) * 5 / 4);
    uint32_t* ro = (uint32_t*)((char*)table + ((size_t)cap << lgp) * 8);
    uint32_t* cursor; cudaMalloc(&cursor, sizeof(uint32_t) << lgp);
    c.lgp = lgp; c.cap = cap; c.cursor = cursor; c.ko = ko; c.ro = ro;
    float ms = timeit(launch_part, &c);
    printf("  partitions 2^%-2d  %7.2f ms  %6.1f G rows/s\n", lgp, ms, n / (ms * 1e6));
    cudaFree(cursor);
  }
  printf("random atomicOr into a bitmap, 2^27 bits set\n");
  for (int lg = 27; lg <= 31; lg += 2) {
    c.mask = (1ull << lg) - 1; c.n = 1ll << 27;
    float ms = timeit(launch_bits, &c);
    printf("  bitmap 2^%d bits (%4d MB)  %7.2f ms  %6.1f G/s\n", lg, (int)((1ull << lg) >> 23), ms, (1 << 27) / (ms * 1e6));
  }
  return 0;
}
