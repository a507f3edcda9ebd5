// Microbenchmark 3: does moving the coalesced key / iKey streams onto the TMA (bulk copies global<->shared) free
// L1TEX request slots for the random table probes?  Shape of the hash-probe round: 8 B/row in, one random 32-byte
// LDG.256 into an L2-resident table, 4 B/row out.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_random3 l2_random3.cu && ./l2_random3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t probe(const void* table, uint32_t idx) {
  unsigned long long a, b, c, d;
  asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"((const uint4*)table + 2 * (size_t)(idx >> 1)));
  return (uint32_t)(a ^ d);
}

constexpr int T = 1024;  // rows per tile
// IN_TMA: key tile arrives by bulk copy; OUT_TMA: result tile leaves by bulk copy.  Contiguous tile runs per CTA.
template <bool IN_TMA, bool OUT_TMA>
__global__ void __launch_bounds__(256) k(const void* table, uint32_t mask, int64_t ntiles, const unsigned long long* in, uint32_t* out) {
  __shared__ __align__(128) unsigned long long keys[2][T];
  __shared__ __align__(128) uint32_t outs[2][T];
  __shared__ unsigned long long bar[2];
  const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
  if (t0 >= t1) return;
  if (IN_TMA) {
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar[0], T * 8); bulk_g2s(keys[0], in + t0 * T, T * 8, &bar[0]); }
  }
  uint32_t ph[2] = {0, 0};
  for (int64_t t = t0; t < t1; t++) {
    const int b = (int)((t - t0) & 1);
    unsigned long long kk[4];
    if (IN_TMA) {
      if (threadIdx.x == 0 && t + 1 < t1) { mbar_expect_tx(&bar[b ^ 1], T * 8); bulk_g2s(keys[b ^ 1], in + (t + 1) * T, T * 8, &bar[b ^ 1]); }
      mbar_wait(&bar[b], ph[b]); ph[b] ^= 1;
#pragma unroll
      for (int j = 0; j < 4; j++) kk[j] = keys[b][j * 256 + threadIdx.x];
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(kk[j]) : "l"(in + t * T + j * 256 + threadIdx.x));
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) o[j] = probe(table, hash32((uint32_t)(t * T + j * 256 + threadIdx.x) ^ (uint32_t)kk[j]) & mask);
    if (OUT_TMA) {
      if (threadIdx.x == 0) bulk_wait_read1();  // the store issued two tiles ago has read its buffer
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; j++) outs[b][j * 256 + threadIdx.x] = o[j];
      fence_proxy_async();
      __syncthreads();
      if (threadIdx.x == 0) bulk_s2g(out + t * T, outs[b], T * 4);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) out[t * T + j * 256 + threadIdx.x] = o[j];
      if (IN_TMA) __syncthreads();  // keys[b] is refilled two iterations later
    }
  }
  if (OUT_TMA && threadIdx.x == 0) bulk_wait_all();
}

template <bool IN_TMA, bool OUT_TMA>
void run(const char* name, void* table, unsigned long long* in, uint32_t* out, int ctas_per_sm) {
  printf("%-44s", name);
  for (int lg = 20; lg <= 22; lg += 1) {  // 16-byte elements: 16 MB .. 64 MB
    uint32_t mask = (1u << lg) - 1;
    int64_t n = 1ll << 29;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<IN_TMA, OUT_TMA><<<148 * ctas_per_sm, 256>>>(table, mask, (n >> 3) / T, in, out);
    cudaEventRecord(a);
    k<IN_TMA, OUT_TMA><<<148 * ctas_per_sm, 256>>>(table, mask, n / T, in, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf(" %6.1f", n / (ms * 1e6));
  }
  cudaError_t e = cudaGetLastError();
  printf("   G rows/s %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  void* table; cudaMalloc(&table, (size_t)1 << 28); cudaMemset(table, 0, (size_t)1 << 28);
  unsigned long long* in; cudaMalloc(&in, (size_t)8 << 29); cudaMemset(in, 1, (size_t)8 << 29);
  uint32_t* out; cudaMalloc(&out, (size_t)4 << 29);
  printf("table bytes:                                   16MB   32MB   64MB\n");
  for (int c = 2; c <= 6; c += 2) {
    char nm[64];
    snprintf(nm, 64, "LDG in, STG out, %d CTAs/SM", c); run<false, false>(nm, table, in, out, c);
    snprintf(nm, 64, "TMA in, STG out, %d CTAs/SM", c); run<true, false>(nm, table, in, out, c);
    snprintf(nm, 64, "TMA in, TMA out, %d CTAs/SM", c); run<true, true>(nm, table, in, out, c);
  }
  return 0;
}
