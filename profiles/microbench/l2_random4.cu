// Microbenchmark 4 (round 2): can the two random accesses of the fused group-and-sum row -- the 32-byte bucket probe
// and the f64 accumulate -- be taken off the LSU / L1TEX t-stage by issuing them through the TMA?
//   A. probe:   one `cp.async.bulk` global->shared of 32 bytes per row (random bucket of an L2-resident table),
//               completion counted in bytes on a per-warp mbarrier, every lane issuing its own copies
//   B. reduce:  one `cp.reduce.async.bulk ... .add.f64` shared->global of 16 bytes per row ({v, 0.0} onto an aligned
//               pair of accumulators), bulk-group completion
//   C. the plain forms beside them: LDG.256 probe, RED.ADD.F64
// Rates are rows (= operations) per second.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart=shared -o l2_random4 l2_random4.cu && ./l2_random4
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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_red_add_f64(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int B = 4;  // operations in flight per thread

// A: TMA probe.  Every lane issues B 32-byte bulk loads per iteration into its own shared-memory slots.
__global__ void __launch_bounds__(256) k_tma_probe(const uint8_t* table, uint32_t bmask, int iters, uint32_t* sink) {
  __shared__ __align__(128) uint8_t slots[256 * B * 32];
  __shared__ unsigned long long bars[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) mbar_init(&bars[w], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t x = hash32(blockIdx.x * 256 + threadIdx.x + 1), acc = 0, parity = 0;
  uint8_t* mine = slots + (size_t)threadIdx.x * B * 32;
  for (int it = 0; it < iters; it++) {
    if (lane == 0) mbar_expect_tx(&bars[w], 32 * B * 32);
    __syncwarp();
#pragma unroll
    for (int b = 0; b < B; b++) {
      x = hash32(x + b);
      bulk_g2s(mine + b * 32, table + (size_t)(x & bmask) * 32, 32, &bars[w]);
    }
    mbar_wait(&bars[w], parity);
    parity ^= 1;
#pragma unroll
    for (int b = 0; b < B; b++) acc ^= *(const uint32_t*)(mine + b * 32);
    __syncwarp();
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// C1: plain LDG.256 probe, B in flight per thread
__global__ void __launch_bounds__(256) k_ldg_probe(const uint8_t* table, uint32_t bmask, int iters, uint32_t* sink) {
  uint32_t x = hash32(blockIdx.x * 256 + threadIdx.x + 1), acc = 0;
  for (int it = 0; it < iters; it++) {
    unsigned long long a[B], d[B];
#pragma unroll
    for (int b = 0; b < B; b++) {
      x = hash32(x + b);
      unsigned long long q, r;
      asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a[b]), "=l"(q), "=l"(r), "=l"(d[b]) : "l"(table + (size_t)(x & bmask) * 32));
    }
#pragma unroll
    for (int b = 0; b < B; b++) acc ^= (uint32_t)(a[b] ^ d[b]);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// B: TMA reduce.  Every lane issues B 16-byte bulk reductions per iteration from its own shared-memory slots.
__global__ void __launch_bounds__(256) k_tma_red(double* accs, uint32_t amask, int iters) {
  __shared__ __align__(128) double src[256 * B * 2];
  uint32_t x = hash32(blockIdx.x * 256 + threadIdx.x + 1);
  double* mine = src + (size_t)threadIdx.x * B * 2;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int b = 0; b < B; b++) {
      mine[2 * b] = 1.0;
      mine[2 * b + 1] = 0.0;
    }
    fence_proxy_async();
#pragma unroll
    for (int b = 0; b < B; b++) {
      x = hash32(x + b);
      bulk_red_add_f64(accs + (size_t)((x & amask) & ~1u), mine + 2 * b, 16);
    }
    bulk_commit();
    bulk_wait_read0();  // the sources may be rewritten
  }
  bulk_wait0();
}

// C2: plain RED.ADD.F64
__global__ void __launch_bounds__(256) k_red(double* accs, uint32_t amask, int iters) {
  uint32_t x = hash32(blockIdx.x * 256 + threadIdx.x + 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int b = 0; b < B; b++) {
      x = hash32(x + b);
      atomicAdd(accs + (x & amask), 1.0);
    }
  }
}

template <typename F>
static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
  return ms;
}

int main() {
  const int iters = 512;
  uint32_t* sink;
  cudaMalloc(&sink, 4);
  printf("# B200: random accesses through the TMA against the plain instructions; G operations / s\n");
  printf("%-28s %8s %8s %8s\n", "table / accumulator bytes:", "8MB", "32MB", "64MB");
  for (int ctas = 2; ctas <= 8; ctas *= 2) {
    const int grid = 148 * ctas;
    const double ops = (double)grid * 256 * B * iters;
    double r[4][3];
    int col = 0;
    for (size_t mb : {8, 32, 64}) {
      const size_t bytes = mb << 20;
      uint8_t* table;
      cudaMalloc(&table, bytes);
      cudaMemset(table, 1, bytes);
      const uint32_t bmask = (uint32_t)(bytes / 32 - 1), amask = (uint32_t)(bytes / 8 - 1);
      double ms;
      ms = time_ms([&] { k_tma_probe<<<grid, 256>>>(table, bmask, iters, sink); });
      r[0][col] = ops / ms / 1e6;
      ms = time_ms([&] { k_ldg_probe<<<grid, 256>>>(table, bmask, iters, sink); });
      r[1][col] = ops / ms / 1e6;
      cudaMemset(table, 0, bytes);
      ms = time_ms([&] { k_tma_red<<<grid, 256>>>((double*)table, amask, iters); });
      r[2][col] = ops / ms / 1e6;
      ms = time_ms([&] { k_red<<<grid, 256>>>((double*)table, amask, iters); });
      r[3][col] = ops / ms / 1e6;
      cudaFree(table);
      col++;
    }
    const char* names[4] = {"TMA 32B probe", "LDG.256 probe", "TMA 16B reduce add.f64", "RED.ADD.F64"};
    for (int k = 0; k < 4; k++) printf("%-22s %dC/SM %8.1f %8.1f %8.1f\n", names[k], ctas, r[k][0], r[k][1], r[k][2]);
  }
  return 0;
}
