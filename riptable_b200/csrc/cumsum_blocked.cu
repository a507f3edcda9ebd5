// a9, rc.EmaAll32 GB_CUMSUM (riptable/rt_groupbyops.py:2664-2724, rt_enum.py:521) without a sort, for groupings of up to
// a few million bins.
//
// The sort-based form (pack.cu) packs the rows by bin, gathers the values into packed order, scans and scatters the
// results back: two row-sized random DRAM passes (a gather at 36 G rows/s, a scatter at 21 G rows/s) around the radix
// passes -- 99 ms per 1 B rows.  A running sum per bin is also a *blocked* scan over the rows in their own order:
//
//   1. totals  M[c][b] = sum of the rows of chunk c that fall in bin b          (one RED per row; the rows of a chunk hit
//                                                                               one (U+1)-sized slice, which stays in L2)
//   2. chunks  M[c][b] <- sum of M[c'][b], c' < c                               (one thread per bin walks the chunks:
//                                                                               coalesced, 2 x the matrix)
//   3. rows    one CTA per chunk walks its rows tile by tile, in order:
//              out[i] = M[c][b_i] + the rows of the chunk before i that fall in b_i, + x_i
//
// Step 3 keeps the running value of every bin a tile touches in a shared-memory table (loaded from / stored back to
// M[c][.], which the CTA owns); the tile's warps take turns in row order, and inside a warp rows of the same bin are
// ordered with match.any, so every bin's values are added in row order.  Everything streams: keys and values are read
// twice, the result written once, the matrix (chunks x bins x 8 bytes, <= 8 GB) crossed three times.
//
// Step 3 touches M[c][b_i] once per row, for every chunk in flight at once: it only pays while the whole matrix stays
// in L2 (measured: 1 M bins x 512 chunks = 4 GB turns every row into a random DRAM read and a random DRAM write, 103 ms
// per 1 B rows -- no better than the sort).  With up to 3 K bins a chunk is owned by a *warp* (thousands of chunks, no
// table and no turns: the running value is read from and written back to M in L2 at every 32-row step); with up to 16 K
// bins by a CTA as described.
//
// Used when there is no reset filter, the bins fit (U + 1 <= 16 Ki) and the call is large enough to fill the GPU;
// everything else stays on the sort-based path.  Integer sums are exact and identical on both paths; float sums are
// accumulated in float64 on both and differ in the association of the partial sums only.
#include "cumsum_blocked.cuh"
#include "values.cuh"

namespace rtb {

constexpr int kCbTile = 1024;          // rows per tile (256 threads x 4)
constexpr int kCbThreads = 256;
constexpr int kCbSlots = 2048;         // shared-memory table slots per tile (<= 50 % full)
constexpr int64_t kCbMinRows = 4ll << 20;
constexpr int64_t kCbMaxBins = 16384;      // CTA-owned chunks: 512 chunks x 16 Ki bins x 8 bytes = 64 MB of L2
constexpr int64_t kCbWarpMaxBins = 3072;   // warp-owned chunks: at least 2048 chunks within 48 MB
constexpr int kCbMaxChunks = 512;
constexpr int kCbMinChunks = 128;
constexpr int kCbWarpChunks = 148 * 8 * 8; // warp-owned: one chunk per resident warp (32 registers: 64 warps per SM)
constexpr size_t kCbMatrixBudget = (size_t)64 << 20;
constexpr size_t kCbWarpMatrixBudget = (size_t)48 << 20;

CbPlan cumsum_blocked_plan(int func, int64_t nrows, int64_t unique_rows, bool has_reset) {
  CbPlan p{};
  // RTB_CUMSUM_BLOCKED: 0 = never (measurement), 2 = whenever it can run at all, however small the call (tests)
  const char* e = getenv("RTB_CUMSUM_BLOCKED");
  const int mode = e ? atoi(e) : 1;
  const bool force = mode == 2;
  const bool served = func == RTB_GB_CUMSUM || func == RTB_GB_CUMMIN || func == RTB_GB_CUMMAX || func == RTB_GB_CUMNANMIN || func == RTB_GB_CUMNANMAX;
  if (mode == 0 || !served || has_reset || nrows < (force ? 1 : kCbMinRows) || unique_rows + 1 > kCbMaxBins) return p;
  p.stride = (unique_rows + 1 + 15) & ~15ll;
  const int64_t tiles = (nrows + kCbTile - 1) / kCbTile;
  if (unique_rows + 1 <= kCbWarpMaxBins) {
    int64_t c = (nrows + 2047) / 2048;  // at least 64 steps per warp
    if (c > kCbWarpChunks) c = kCbWarpChunks;
    const int64_t cmem = (int64_t)(kCbWarpMatrixBudget / ((size_t)p.stride * 8));
    if (c > cmem) c = cmem;
    if (c >= 1 && (force || c >= 2048)) {
      int64_t rows = (nrows + c - 1) / c;
      p.chunk_rows = (rows + 255) & ~255ll;
      p.chunks = (int)((nrows + p.chunk_rows - 1) / p.chunk_rows);
      p.warp_chunks = true;
      p.bytes = (size_t)p.chunks * (size_t)p.stride * 8 + 16;
      p.ok = true;
      return p;
    }
  }
  int64_t cmem = (int64_t)(kCbMatrixBudget / ((size_t)p.stride * 8));
  int64_t c = (nrows + 16 * kCbTile - 1) / (16 * kCbTile);
  if (c > kCbMaxChunks) c = kCbMaxChunks;
  if (c > cmem) c = cmem;
  if (c < (force ? 1 : kCbMinChunks)) return p;
  int64_t tpc = (tiles + c - 1) / c;
  p.chunk_rows = tpc * kCbTile;
  p.chunks = (int)((nrows + p.chunk_rows - 1) / p.chunk_rows);
  p.bytes = (size_t)p.chunks * (size_t)p.stride * 8 + 16;
  p.ok = true;
  return p;
}

// OP: what a bin accumulates.  Sums run on int64 / float64 bit patterns; running minima / maxima run on the order-preserving
// codes of values.cuh, whose two extreme codes are free: one is the identity ("nothing seen yet"), the other absorbs
// (skipna = 0: a NaN / sentinel poisons the rest of the group), and both come out as the invalid value.
enum { CB_SUM_I = 0, CB_SUM_F = 1, CB_MIN = 2, CB_MAX = 3 };

template <int OP>
__device__ __forceinline__ unsigned long long cb_identity() {
  return OP == CB_SUM_F ? (unsigned long long)__double_as_longlong(0.0) : OP == CB_MIN ? ~0ull : 0ull;
}
template <int OP>
__device__ __forceinline__ unsigned long long cb_poison() {
  return OP == CB_MIN ? 0ull : ~0ull;
}
template <int OP>
__device__ __forceinline__ unsigned long long cb_value(unsigned long long raw, int vdt, int skipna) {
  if (OP == CB_SUM_F) return (unsigned long long)__double_as_longlong(raw_to_f64(raw, vdt));
  if (OP == CB_SUM_I) return (unsigned long long)raw_to_i64(raw, vdt);
  if (raw_invalid(raw, vdt)) return skipna ? cb_identity<OP>() : cb_poison<OP>();
  return raw_to_code(raw, vdt, OP == CB_MAX);
}
template <int OP>
__device__ __forceinline__ unsigned long long cb_add(unsigned long long a, unsigned long long b) {
  if (OP == CB_SUM_F) return (unsigned long long)__double_as_longlong(__longlong_as_double((long long)a) + __longlong_as_double((long long)b));
  if (OP == CB_MIN) return a < b ? a : b;
  if (OP == CB_MAX) return a > b ? a : b;
  return a + b;
}
template <int OP>
__device__ __forceinline__ void cb_atomic(unsigned long long* p, unsigned long long v) {
  if (OP == CB_SUM_F) atomicAdd((double*)p, __longlong_as_double((long long)v));
  else if (OP == CB_MIN) atomicMin(p, v);
  else if (OP == CB_MAX) atomicMax(p, v);
  else atomicAdd(p, v);
}
// the stored form of a running value (bins > 0 only)
template <int OP>
__device__ __forceinline__ unsigned long long cb_out(unsigned long long acc, int vdt, int odt) {
  if (OP == CB_MIN || OP == CB_MAX) return (acc == cb_identity<OP>() || acc == cb_poison<OP>()) ? invalid_raw(odt) : code_to_raw(acc, vdt, OP == CB_MAX);
  if (odt == RTB_F32) return (unsigned long long)__float_as_uint((float)__longlong_as_double((long long)acc));
  return acc;
}

// ---- 1. per-chunk totals ----------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256) cb_totals_kernel(const void* __restrict__ values, int vdt, int skipna, const void* __restrict__ ikey, int kdt,
                                                        int64_t n, int64_t unique_rows, const uint8_t* __restrict__ filter,
                                                        int64_t chunk_rows, int64_t stride, unsigned long long* __restrict__ M,
                                                        uint32_t* bad) {
  const int vsz = dsize(vdt);
  const int64_t ntiles = (n + kCbTile - 1) / kCbTile;
  uint32_t nbad = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t lo = t * kCbTile;
    unsigned long long* row = M + (lo / chunk_rows) * stride;
#pragma unroll
    for (int k = 0; k < kCbTile / kCbThreads; k++) {
      const int64_t i = lo + k * kCbThreads + threadIdx.x;
      if (i >= n) break;
      const int64_t b = load_ikey(ikey, kdt, i);
      if (b < 0 || b > unique_rows) nbad++;
      if (b <= 0 || b > unique_rows || (filter && !filter[i])) continue;
      const unsigned long long v = cb_value<OP>(load_raw1(values, vsz, i), vdt, skipna);
      cb_atomic<OP>(&row[b], v);
    }
  }
  if (nbad) atomicAdd(bad, nbad);  // rows outside [0, unique_rows]: the caller rejects the input, as the packed form does
}

// Few bins: one RED per row would queue on a handful of L2 addresses.  A CTA sums a contiguous part of one chunk in shared
// memory instead (rows of the same bin inside a warp are combined first) and adds its part to the chunk's slice once.
constexpr int kCbSmemBins = 4096;
constexpr int kCbParts = 8;  // CTAs per chunk
template <int OP>
__global__ void __launch_bounds__(256) cb_totals_smem_kernel(const void* __restrict__ values, int vdt, int skipna, const void* __restrict__ ikey, int kdt,
                                                             int64_t n, int64_t unique_rows, const uint8_t* __restrict__ filter,
                                                             int64_t chunk_rows, int64_t stride, unsigned long long* __restrict__ M,
                                                             uint32_t* bad) {
  __shared__ unsigned long long acc[kCbSmemBins];
  uint32_t nbad = 0;
  const int nb = (int)unique_rows + 1;
  const unsigned long long zero = cb_identity<OP>();
  for (int b = threadIdx.x; b < nb; b += blockDim.x) acc[b] = zero;
  __syncthreads();
  const int vsz = dsize(vdt);
  const int lane = threadIdx.x & 31;
  const int64_t clo = (int64_t)blockIdx.y * chunk_rows;
  const int64_t chi = clo + chunk_rows < n ? clo + chunk_rows : n;
  const int64_t part = ((chunk_rows / gridDim.x) + 255) & ~255ll;
  const int64_t lo = clo + (int64_t)blockIdx.x * part;
  const int64_t hi = lo + part < chi ? lo + part : chi;
  for (int64_t base = lo; base < hi; base += blockDim.x) {  // base is the same for the whole CTA: full-warp collectives
    const int64_t i = base + threadIdx.x;
    int b = -1;
    unsigned long long v = zero;
    if (i < hi) {
      const int64_t bb = load_ikey(ikey, kdt, i);
      if (bb < 0 || bb > unique_rows) nbad++;
      if (bb > 0 && bb <= unique_rows && (!filter || filter[i])) {
        b = (int)bb;
        v = cb_value<OP>(load_raw1(values, vsz, i), vdt, skipna);
      }
    }
    const uint32_t m = __match_any_sync(0xffffffffu, b);
    uint32_t lower = b >= 0 ? (m & ((1u << lane) - 1u)) : 0u;
    unsigned long long sum = v;
    while (__any_sync(0xffffffffu, lower != 0)) {
      const int src = lower ? (__ffs(lower) - 1) : lane;
      const unsigned long long t = __shfl_sync(0xffffffffu, v, src);
      if (lower) {
        sum = cb_add<OP>(sum, t);
        lower &= lower - 1;
      }
    }
    if (b >= 0 && (m >> lane) == 1u) {  // the last lane of a bin holds the sum of the bin's rows in this warp
      cb_atomic<OP>(&acc[b], sum);
    }
  }
  if (nbad) atomicAdd(bad, nbad);
  __syncthreads();
  unsigned long long* row = M + (int64_t)blockIdx.y * stride;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    const unsigned long long a = acc[b];
    if (a != zero) {
      cb_atomic<OP>(&row[b], a);
    }
  }
}

// ---- 2. exclusive scan over the chunks, one thread per bin ---------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256) cb_scan_chunks_kernel(unsigned long long* __restrict__ M, int chunks, int64_t stride, int64_t nbins) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  unsigned long long run = cb_identity<OP>();
  unsigned long long* p = M + b;
  int c = 0;
  for (; c + 8 <= chunks; c += 8) {  // the loads of a batch are independent of the running sum
    unsigned long long t[8];
#pragma unroll
    for (int j = 0; j < 8; j++) t[j] = p[(int64_t)(c + j) * stride];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      p[(int64_t)(c + j) * stride] = run;
      run = cb_add<OP>(run, t[j]);
    }
  }
  for (; c < chunks; c++) {
    const unsigned long long t = p[(int64_t)c * stride];
    p[(int64_t)c * stride] = run;
    run = cb_add<OP>(run, t);
  }
}

// Few bins, thousands of chunks (the warp-owned form): a warp per bin, 32 chunks per step, instead of one thread walking
// 9 472 chunks.
template <int OP>
__global__ void __launch_bounds__(256) cb_scan_chunks_warp_kernel(unsigned long long* __restrict__ M, int chunks, int64_t stride, int64_t nbins) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= nbins) return;
  unsigned long long run = cb_identity<OP>();
  unsigned long long* p = M + b;
  for (int c0 = 0; c0 < chunks; c0 += 32) {
    const int c = c0 + lane;
    const unsigned long long t = c < chunks ? p[(int64_t)c * stride] : cb_identity<OP>();
    unsigned long long inc = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc = cb_add<OP>(u, inc);
    }
    unsigned long long ex = __shfl_up_sync(0xffffffffu, inc, 1);
    ex = lane ? cb_add<OP>(run, ex) : run;
    if (c < chunks) p[(int64_t)c * stride] = ex;
    run = cb_add<OP>(run, __shfl_sync(0xffffffffu, inc, 31));
  }
}

// ---- 3. the rows of a chunk, in order --------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cb_ld_cg(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void cb_st_cg(unsigned long long* p, unsigned long long v) {
  asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// A tile's rows fall into two classes.  A row whose bin occurs once in the tile needs no ordering at all: it adds its value
// to the bin's running value, all such rows at once.  Rows whose bin occurs again in the tile (about 1024^2 / 2U of them:
// 128 at 4 K bins, 32 at 16 K) are compacted in row order into a short list, which one warp walks 32 at a time, ordering
// rows of the same bin with match.any.  (Taking the whole tile through that ordered walk -- the first version -- cost 32
// serial steps per tile, 17 us; with few bins, where most rows repeat, the warp-owned form is used instead.)
template <int OP>
__global__ void __launch_bounds__(kCbThreads, 4) cb_rows_kernel(const void* __restrict__ values, int vdt, int skipna, int odt,
                                                              const void* __restrict__ ikey, int kdt, int64_t n, int64_t unique_rows,
                                                              const uint8_t* __restrict__ filter, int64_t chunk_rows, int64_t stride,
                                                              unsigned long long* __restrict__ M, void* __restrict__ out) {
  __shared__ int32_t skey[kCbSlots];             // bin held by the slot, -1 = free
  __shared__ uint32_t scnt[kCbSlots];            // rows of the tile that fall in it
  __shared__ unsigned long long sval[kCbSlots];  // its running value
  __shared__ int32_t lslot[kCbTile];             // the repeated rows, in row order: slot ...
  __shared__ unsigned long long lval[kCbTile];   // ... and value, replaced by the row's result
  __shared__ uint32_t wtot[kCbThreads / 32 + 1];
  constexpr int R = kCbTile / kCbThreads;        // rows per thread; a warp owns 32 * R consecutive rows of the tile
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int vsz = dsize(vdt);
  const int64_t lo = (int64_t)blockIdx.x * chunk_rows;
  const int64_t hi = lo + chunk_rows < n ? lo + chunk_rows : n;
  unsigned long long* row = M + (int64_t)blockIdx.x * stride;
  const unsigned long long inv = invalid_raw(odt);
  const unsigned long long zero = cb_identity<OP>();

  // the next tile's keys / values / filter bytes are in flight while this tile is resolved
  int64_t nbin[R];
  unsigned long long nraw[R];
  uint8_t nflt[R];
  auto fetch = [&](int64_t tlo) {
#pragma unroll
    for (int k = 0; k < R; k++) {
      const int64_t i = tlo + w * (32 * R) + k * 32 + lane;
      nbin[k] = 0;
      nraw[k] = 0;
      nflt[k] = 1;
      if (i < hi) {
        nbin[k] = load_ikey(ikey, kdt, i);
        nraw[k] = load_raw1(values, vsz, i);
        if (filter) nflt[k] = filter[i];
      }
    }
  };
  fetch(lo);
  for (int64_t tlo = lo; tlo < hi; tlo += kCbTile) {
    int32_t bin[R];
    unsigned long long val[R];
    int slot[R];
    bool owner[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      const bool live = nbin[k] > 0 && nbin[k] <= unique_rows;
      bin[k] = live ? (int32_t)nbin[k] : 0;
      val[k] = (live && nflt[k]) ? cb_value<OP>(nraw[k], vdt, skipna) : zero;
    }
    if (tlo + kCbTile < hi) fetch(tlo + kCbTile);
#pragma unroll
    for (int j = 0; j < kCbSlots / kCbThreads; j++) {
      skey[j * kCbThreads + tid] = -1;
      scnt[j * kCbThreads + tid] = 0;
    }
    __syncthreads();
    // every distinct bin of the tile gets a slot; the row that claims it loads the bin's running value
#pragma unroll
    for (int k = 0; k < R; k++) {
      slot[k] = -1;
      owner[k] = false;
      if (bin[k]) {
        uint32_t h = ((uint32_t)bin[k] * 0x9E3779B1u) >> (32 - 11);
        for (;;) {
          const int32_t old = atomicCAS(&skey[h], -1, bin[k]);
          if (old == -1) { owner[k] = true; break; }
          if (old == bin[k]) break;
          h = (h + 1) & (kCbSlots - 1);
        }
        slot[k] = (int)h;
        atomicAdd(&scnt[h], 1u);
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++)
      if (owner[k]) sval[slot[k]] = cb_ld_cg(&row[bin[k]]);
    __syncthreads();
    // rows whose bin repeats in the tile go to the ordered list (positions by an exclusive scan of the flags in row order:
    // warp w holds rows w*128 .. w*128+127 as R steps of 32)
    bool rep[R];
    uint32_t before[R];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < R; k++) {
      rep[k] = slot[k] >= 0 && scnt[slot[k]] > 1u;
      const uint32_t b = __ballot_sync(0xffffffffu, rep[k]);
      before[k] = mine + __popc(b & ((1u << lane) - 1u));
      mine += __popc(b);
    }
    if (lane == 0) wtot[w] = mine;
    __syncthreads();
    uint32_t base = 0, nrep = 0;
#pragma unroll
    for (int j = 0; j < kCbThreads / 32; j++) {
      if (j < w) base += wtot[j];
      nrep += wtot[j];
    }
    unsigned long long res[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      res[k] = zero;
      if (rep[k]) {
        lslot[base + before[k]] = slot[k];
        lval[base + before[k]] = val[k];
      } else if (slot[k] >= 0) {  // the bin's only row in this tile
        res[k] = cb_add<OP>(sval[slot[k]], val[k]);
        sval[slot[k]] = res[k];
      }
    }
    __syncthreads();
    if (w == 0) {  // one warp walks the repeated rows in row order
      for (uint32_t j0 = 0; j0 < nrep; j0 += 32) {
        const uint32_t j = j0 + lane;
        const int s_ = j < nrep ? lslot[j] : -1;
        const unsigned long long v = j < nrep ? lval[j] : zero;
        const uint32_t m = __match_any_sync(0xffffffffu, s_);
        unsigned long long acc = s_ >= 0 ? sval[s_] : zero;
        uint32_t lower = s_ >= 0 ? (m & ((1u << lane) - 1u)) : 0u;
        while (__any_sync(0xffffffffu, lower != 0)) {
          const int src = lower ? (__ffs(lower) - 1) : lane;
          const unsigned long long t = __shfl_sync(0xffffffffu, v, src);
          if (lower) {
            acc = cb_add<OP>(acc, t);
            lower &= lower - 1;
          }
        }
        acc = cb_add<OP>(acc, v);
        if (s_ >= 0) {
          lval[j] = acc;
          if ((m >> lane) == 1u) sval[s_] = acc;  // the last row of the bin in this step
        }
        __syncwarp();
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < R; k++) {
      const int64_t i = tlo + w * (32 * R) + k * 32 + lane;
      if (rep[k]) res[k] = lval[base + before[k]];
      if (i < hi) {
        unsigned long long o = inv;
        if (bin[k]) o = cb_out<OP>(res[k], vdt, odt);
        store_raw(out, odt, i, o);
      }
      if (owner[k]) cb_st_cg(&row[bin[k]], sval[slot[k]]);
    }
    __syncthreads();  // the stored running values are visible to the loads of the next tile; the tables may be cleared
  }
}

// Few bins: one warp per chunk, 32 rows per step.  Rows of the same bin inside the step are ordered with match.any; the
// bin's running value lives in M[c][b] (L2), read by every row of the bin and written back by the last one.
template <int OP>
__global__ void __launch_bounds__(256) cb_rows_warp_kernel(const void* __restrict__ values, int vdt, int skipna, int odt,
                                                           const void* __restrict__ ikey, int kdt, int64_t n, int64_t unique_rows,
                                                           const uint8_t* __restrict__ filter, int64_t chunk_rows, int64_t stride,
                                                           int chunks, unsigned long long* __restrict__ M, void* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= chunks) return;
  const int vsz = dsize(vdt);
  const int64_t lo = c * chunk_rows;
  const int64_t hi = lo + chunk_rows < n ? lo + chunk_rows : n;
  unsigned long long* row = M + c * stride;
  const unsigned long long inv = invalid_raw(odt);
  const unsigned long long zero = cb_identity<OP>();
  // the next step's key / value / filter are in flight while this step is resolved
  int64_t nb = 0;
  unsigned long long nraw = 0;
  uint8_t nf = 1;
  if (lo + lane < hi) {
    nb = load_ikey(ikey, kdt, lo + lane);
    nraw = load_raw1(values, vsz, lo + lane);
    if (filter) nf = filter[lo + lane];
  }
  for (int64_t base = lo; base < hi; base += 32) {
    const int64_t i = base + lane;
    const int64_t b = nb;
    const unsigned long long raw = nraw;
    const uint8_t f = nf;
    nb = 0;
    if (i + 32 < hi) {
      nb = load_ikey(ikey, kdt, i + 32);
      nraw = load_raw1(values, vsz, i + 32);
      if (filter) nf = filter[i + 32];
    }
    const bool live = i < hi && b > 0 && b <= unique_rows;
    const int key = live ? (int)b : -1;
    const unsigned long long v = (live && f) ? cb_value<OP>(raw, vdt, skipna) : zero;
    const uint32_t m = __match_any_sync(0xffffffffu, key);
    unsigned long long acc = live ? cb_ld_cg(&row[b]) : zero;
    uint32_t lower = live ? (m & ((1u << lane) - 1u)) : 0u;
    while (__any_sync(0xffffffffu, lower != 0)) {
      const int src = lower ? (__ffs(lower) - 1) : lane;
      const unsigned long long t = __shfl_sync(0xffffffffu, v, src);
      if (lower) {
        acc = cb_add<OP>(acc, t);
        lower &= lower - 1;
      }
    }
    acc = cb_add<OP>(acc, v);
    if (live && (m >> lane) == 1u) cb_st_cg(&row[b], acc);
    if (i < hi) {
      unsigned long long o = inv;
      if (live) o = cb_out<OP>(acc, vdt, odt);
      store_raw(out, odt, i, o);
    }
    __syncwarp();  // the stored running values are ordered before the next step's loads
  }
}

int cumsum_blocked_run(const CbPlan& p, int func, const void* values, int vdt, int odt, const void* ikey, int kdt, int64_t nrows,
                       int64_t unique_rows, const uint8_t* filter, void* out, void* workspace, cudaStream_t stream) {
  unsigned long long* M = (unsigned long long*)workspace;
  uint32_t* bad = (uint32_t*)((unsigned char*)workspace + p.bytes - 16);  // the plan's bytes end with a 16-byte counter slot
  const bool fl = vdt == RTB_F32 || vdt == RTB_F64;
  const bool ismin = func == RTB_GB_CUMMIN || func == RTB_GB_CUMNANMIN, ismax = func == RTB_GB_CUMMAX || func == RTB_GB_CUMNANMAX;
  const int skipna = func == RTB_GB_CUMNANMIN || func == RTB_GB_CUMNANMAX;
  // the accumulators start at the operation's identity: all-ones bytes for a running minimum, zero bytes otherwise
  RTB_CUDA(cudaMemsetAsync(M, ismin ? 0xFF : 0, p.bytes - 16, stream));
  RTB_CUDA(cudaMemsetAsync(bad, 0, 16, stream));
  const int64_t ntiles = (nrows + kCbTile - 1) / kCbTile;
  const int g1 = (int)(ntiles < (int64_t)kSMs * 8 ? ntiles : (int64_t)kSMs * 8);
  const int g2 = (int)((unique_rows + 1 + 255) / 256);
  const bool few = unique_rows + 1 <= kCbSmemBins;
  int parts = (int)(p.chunk_rows / 16384);
  parts = parts < 1 ? 1 : (parts > kCbParts ? kCbParts : parts);
  const dim3 gs((unsigned)parts, (unsigned)p.chunks);
  const int gw = (p.chunks + 7) / 8;
#define RTB_CB_RUN(FL_)                                                                                                              \
  do {                                                                                                                               \
    if (few) cb_totals_smem_kernel<FL_><<<gs, 256, 0, stream>>>(values, vdt, skipna, ikey, kdt, nrows, unique_rows, filter, p.chunk_rows, p.stride, M, bad); \
    else cb_totals_kernel<FL_><<<g1, 256, 0, stream>>>(values, vdt, skipna, ikey, kdt, nrows, unique_rows, filter, p.chunk_rows, p.stride, M, bad); \
    RTB_LAUNCH_CHECK();                                                                                                              \
    if (p.warp_chunks && unique_rows + 1 <= 4096)                                                                                    \
      cb_scan_chunks_warp_kernel<FL_><<<(int)((unique_rows + 1 + 7) / 8), 256, 0, stream>>>(M, p.chunks, p.stride, unique_rows + 1); \
    else                                                                                                                             \
      cb_scan_chunks_kernel<FL_><<<g2, 256, 0, stream>>>(M, p.chunks, p.stride, unique_rows + 1);                                     \
    RTB_LAUNCH_CHECK();                                                                                                              \
    if (p.warp_chunks)                                                                                                               \
      cb_rows_warp_kernel<FL_><<<gw, 256, 0, stream>>>(values, vdt, skipna, odt, ikey, kdt, nrows, unique_rows, filter, p.chunk_rows, p.stride, \
                                                      p.chunks, M, out);                                                             \
    else                                                                                                                             \
      cb_rows_kernel<FL_><<<p.chunks, kCbThreads, 0, stream>>>(values, vdt, skipna, odt, ikey, kdt, nrows, unique_rows, filter, p.chunk_rows, \
                                                              p.stride, M, out);                                                     \
  } while (0)
  if (ismin) RTB_CB_RUN(CB_MIN);
  else if (ismax) RTB_CB_RUN(CB_MAX);
  else if (fl) RTB_CB_RUN(CB_SUM_F);
  else RTB_CB_RUN(CB_SUM_I);
#undef RTB_CB_RUN
  RTB_LAUNCH_CHECK();
  uint32_t hbad = 0;
  RTB_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  RTB_CUDA(cudaStreamSynchronize(stream));
  RTB_ARG(hbad == 0, "EmaAll32: %u ikey values lie outside [0, %lld]", hbad, (long long)unique_rows);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

}  // namespace rtb
