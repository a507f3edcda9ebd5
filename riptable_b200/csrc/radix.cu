// Stable LSD radix sort engine (see radix.cuh).
#include "radix.cuh"

namespace rtb {

RadixLayout radix_layout(void* ws, size_t wsbytes, int64_t n) {
  Arena a(ws, wsbytes);
  RadixLayout L;
  int64_t tiles = (n + kRadixTile - 1) / kRadixTile;
  if (tiles < 1) tiles = 1;
  int64_t nch = tiles < (int64_t)kSMs * 4 ? tiles : (int64_t)kSMs * 4;
  int64_t tiles_per = (tiles + nch - 1) / nch;
  L.chunk_rows = tiles_per * kRadixTile;
  int64_t real = (n + L.chunk_rows - 1) / L.chunk_rows;
  L.nchunks = (int)(real < 1 ? 1 : real);
  L.counts = a.take<uint32_t>((size_t)256 * L.nchunks);
  L.scan_scratch = a.take<uint32_t>(scan_u32_workspace_elems((int64_t)256 * L.nchunks));
  L.bits = a.take<unsigned long long>(2);
  L.bytes = a.off;
  return L;
}

// ---- which key bits vary at all: OR / AND over every key ------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(256) radix_key_bits(const K* __restrict__ keys, int64_t n, unsigned long long* bits) {
  unsigned long long o = 0, a = ~0ull;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long k = (unsigned long long)keys[i];
    o |= k;
    a &= k;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    o |= __shfl_xor_sync(0xffffffffu, o, d);
    a &= __shfl_xor_sync(0xffffffffu, a, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(&bits[0], o);
    atomicAnd(&bits[1], a);
  }
}

// ---- per-chunk digit counts ---------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(256) radix_chunk_hist(const K* __restrict__ keys, int64_t n, int shift,
                                                        int64_t chunk_rows, int nchunks, uint32_t* __restrict__ counts) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * chunk_rows;
  const int64_t hi = lo + chunk_rows < n ? lo + chunk_rows : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & 0xFFu], 1u);
  __syncthreads();
  counts[(size_t)threadIdx.x * nchunks + blockIdx.x] = sh[threadIdx.x];
}

// ---- scatter ----------------------------------------------------------------------------------------
// exclusive scan of one value per thread across the 256-thread block
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* warp_sums /* 8 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int i = 0; i < 8; i++)
    if (i < w) base += warp_sums[i];
  return base + inc - v;
}

// ---- TMA / mbarrier helpers (bulk 1-D copies global -> shared, completion counted in bytes on an mbarrier) -------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename K>
constexpr size_t radix_scatter_smem() {
  return (size_t)2 * kRadixTile * (sizeof(K) + 4) + (8 * 256 + 3 * 256 + 8) * 4 + 2 * 8 + 16;
}

// One CTA walks the tiles of its chunk.  Tiles are double-buffered in shared memory: while tile t is ranked and
// written out, the bulk copy (TMA, `cp.async.bulk` completing on an mbarrier) of tile t+1 is in flight, so the HBM
// latency of the tile loads — what the register-staged version of this kernel spent most of its time waiting on —
// is hidden.  The buffer a tile arrived in doubles as its digit-ordered staging area before the global stores.
template <typename K, bool IOTA>
__global__ void __launch_bounds__(kRadixThreads, 2)
radix_scatter(const K* __restrict__ kin, const uint32_t* __restrict__ vin, K* __restrict__ kout,
              uint32_t* __restrict__ vout, int64_t n, int shift, int64_t chunk_rows, int nchunks,
              const uint32_t* __restrict__ offsets) {
  extern __shared__ __align__(128) unsigned char smem[];
  K* kbuf = (K*)smem;                                   // [2][tile]
  uint32_t* vbuf = (uint32_t*)(kbuf + 2 * kRadixTile);  // [2][tile]
  uint32_t* wcnt = vbuf + 2 * kRadixTile;               // [8 warps][256 digits]: running count, then exclusive warp offset
  uint32_t* dbase = wcnt + 8 * 256;                     // first tile position of each digit
  uint32_t* goff = dbase + 256;                         // next global output position of each digit for this chunk
  uint32_t* tcnt = goff + 256;                          // rows of this tile per digit
  uint32_t* wsum = tcnt + 256;                          // 8
  unsigned long long* bar = (unsigned long long*)(wsum + 8 + 2);  // 2 mbarriers (8-byte aligned)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t lo = (int64_t)blockIdx.x * chunk_rows;
  const int64_t hi = lo + chunk_rows < n ? lo + chunk_rows : n;
  const int ntiles = (int)((hi - lo + kRadixTile - 1) / kRadixTile);
  goff[tid] = offsets[(size_t)tid * nchunks + blockIdx.x];
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t) {  // called by thread 0: start the bulk copy of a FULL tile t into buffer t & 1
    const int64_t tlo = lo + (int64_t)t * kRadixTile;
    if (t < ntiles && tlo + kRadixTile <= hi) {
      const int b = t & 1;
      mbar_expect_tx(&bar[b], (uint32_t)(kRadixTile * (sizeof(K) + (IOTA ? 0 : 4))));
      bulk_g2s(kbuf + b * kRadixTile, kin + tlo, (uint32_t)(kRadixTile * sizeof(K)), &bar[b]);
      if (!IOTA) bulk_g2s(vbuf + b * kRadixTile, vin + tlo, (uint32_t)(kRadixTile * 4), &bar[b]);
    }
  };
  if (tid == 0) {
    issue(0);
    issue(1);
  }

  for (int t = 0; t < ntiles; t++) {
    const int64_t tile_lo = lo + (int64_t)t * kRadixTile;
    const int b = t & 1;
    K* skeys = kbuf + b * kRadixTile;
    uint32_t* svals = vbuf + b * kRadixTile;
    const bool full = tile_lo + kRadixTile <= hi;
#pragma unroll
    for (int j = 0; j < 8; j++) wcnt[j * 256 + tid] = 0;

    K key[kRadixItems];
    uint32_t val[kRadixItems];
    uint32_t rank[kRadixItems];
    const int wofs = w * 32 * kRadixItems + lane;
    const int64_t wbase = tile_lo + wofs;
    if (full) {
      mbar_wait(&bar[b], (uint32_t)((t >> 1) & 1));
#pragma unroll
      for (int i = 0; i < kRadixItems; i++) {
        key[i] = skeys[wofs + i * 32];
        val[i] = IOTA ? (uint32_t)(wbase + i * 32) : svals[wofs + i * 32];
      }
    } else {
#pragma unroll
      for (int i = 0; i < kRadixItems; i++) {
        int64_t idx = wbase + i * 32;
        bool valid = idx < hi;
        key[i] = valid ? kin[idx] : (K)0;
        val[i] = IOTA ? (uint32_t)idx : (valid ? vin[idx] : 0u);
      }
    }
    __syncthreads();  // every thread holds its rows: the buffer is free to stage the output; wcnt is zeroed
    // stable rank inside the warp's 512 rows: rows are visited in memory order (i major, lane minor)
#pragma unroll
    for (int i = 0; i < kRadixItems; i++) {
      bool valid = wbase + i * 32 < hi;
      uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & 0xFFu) : 256u;
      uint32_t m = __match_any_sync(0xffffffffu, d);
      int leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (lane == leader && valid) {
        old = wcnt[w * 256 + d];
        wcnt[w * 256 + d] = old + __popc(m);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[i] = old + __popc(m & ((1u << lane) - 1u));
      __syncwarp();
    }
    __syncthreads();
    {  // thread = digit: per-warp exclusive offsets, tile count, tile-exclusive digit base
      uint32_t run = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint32_t c = wcnt[j * 256 + tid];
        wcnt[j * 256 + tid] = run;
        run += c;
      }
      tcnt[tid] = run;
      dbase[tid] = block_excl_scan_256(run, wsum);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRadixItems; i++) {
      if (wbase + i * 32 < hi) {
        uint32_t d = (uint32_t)(key[i] >> shift) & 0xFFu;
        uint32_t p = dbase[d] + wcnt[w * 256 + d] + rank[i];
        skeys[p] = key[i];
        svals[p] = val[i];
      }
    }
    __syncthreads();
    const int tile_n = (int)(hi - tile_lo < kRadixTile ? hi - tile_lo : kRadixTile);
    for (int p = tid; p < tile_n; p += kRadixThreads) {
      K k = skeys[p];
      uint32_t d = (uint32_t)(k >> shift) & 0xFFu;
      uint32_t g = goff[d] + ((uint32_t)p - dbase[d]);
      if (kout) kout[g] = k;
      vout[g] = svals[p];
    }
    fence_proxy_async();  // this thread's generic accesses to the buffer precede the next bulk copy into it
    __syncthreads();
    goff[tid] += tcnt[tid];
    if (tid == 0) issue(t + 2);
  }
}

__global__ void __launch_bounds__(256) fill_iota_u32(uint32_t* out, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i;
}

template <typename K>
int radix_sort_pairs(K* ka, uint32_t* va, K* kb, uint32_t* vb, int64_t n, int key_bits, bool va_iota, void* ws,
                     size_t wsbytes, cudaStream_t stream, int* result_in_b, int* passes_run) {
  *result_in_b = 0;
  if (passes_run) *passes_run = 0;
  if (n <= 0) return RTB_OK;
  RadixLayout L = radix_layout(ws, wsbytes, n);
  RTB_ARG(ws && L.bytes <= wsbytes, "radix sort: workspace too small (%zu < %zu)", wsbytes, L.bytes);
  static bool attr_done = false;
  if (!attr_done) {
    RTB_CUDA(cudaFuncSetAttribute(radix_scatter<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)radix_scatter_smem<uint32_t>()));
    RTB_CUDA(cudaFuncSetAttribute(radix_scatter<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)radix_scatter_smem<uint32_t>()));
    RTB_CUDA(cudaFuncSetAttribute(radix_scatter<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)radix_scatter_smem<uint64_t>()));
    RTB_CUDA(cudaFuncSetAttribute(radix_scatter<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)radix_scatter_smem<uint64_t>()));
    attr_done = true;
  }
  // which bits vary
  unsigned long long hbits[2] = {0ull, ~0ull};
  RTB_CUDA(cudaMemcpyAsync(L.bits, hbits, sizeof(hbits), cudaMemcpyHostToDevice, stream));
  radix_key_bits<K><<<grid_for(n, 256, 8, 8), 256, 0, stream>>>(ka, n, L.bits);
  RTB_LAUNCH_CHECK();
  RTB_CUDA(cudaMemcpyAsync(hbits, L.bits, sizeof(hbits), cudaMemcpyDeviceToHost, stream));
  RTB_CUDA(cudaStreamSynchronize(stream));
  unsigned long long varying = hbits[0] ^ hbits[1];
  if (key_bits < 64) varying &= (1ull << key_bits) - 1ull;

  K* kin = ka;
  K* kout = kb;
  uint32_t* vin = va;
  uint32_t* vout = vb;
  bool iota = va_iota;
  const size_t smem = radix_scatter_smem<K>();
  for (int shift = 0; shift < 8 * (int)sizeof(K) && shift < key_bits; shift += 8) {
    if (((varying >> shift) & 0xFFull) == 0) continue;
    prof_begin(PROF_RADIX_HIST, stream);
    radix_chunk_hist<K><<<L.nchunks, 256, 0, stream>>>(kin, n, shift, L.chunk_rows, L.nchunks, L.counts);
    RTB_LAUNCH_CHECK();
    prof_end(PROF_RADIX_HIST, stream, (uint64_t)n);
    int rc = exclusive_scan_u32(L.counts, L.counts, (int64_t)256 * L.nchunks, SCAN_IDENTITY, L.scan_scratch, nullptr, stream);
    if (rc != RTB_OK) return rc;
    prof_begin(PROF_RADIX_SCATTER, stream);
    if (iota)
      radix_scatter<K, true><<<L.nchunks, kRadixThreads, smem, stream>>>(kin, vin, kout, vout, n, shift, L.chunk_rows, L.nchunks, L.counts);
    else
      radix_scatter<K, false><<<L.nchunks, kRadixThreads, smem, stream>>>(kin, vin, kout, vout, n, shift, L.chunk_rows, L.nchunks, L.counts);
    RTB_LAUNCH_CHECK();
    prof_end(PROF_RADIX_SCATTER, stream, (uint64_t)n);
    K* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
    iota = false;
    *result_in_b ^= 1;
    if (passes_run) ++*passes_run;
  }
  if (iota) {  // no digit varied: the identity permutation
    fill_iota_u32<<<grid_for(n, 256, 4), 256, 0, stream>>>(va, n);
    RTB_LAUNCH_CHECK();
  }
  return RTB_OK;
}

template int radix_sort_pairs<uint32_t>(uint32_t*, uint32_t*, uint32_t*, uint32_t*, int64_t, int, bool, void*, size_t, cudaStream_t, int*, int*);
template int radix_sort_pairs<uint64_t>(uint64_t*, uint32_t*, uint64_t*, uint32_t*, int64_t, int, bool, void*, size_t, cudaStream_t, int*, int*);

}  // namespace rtb
