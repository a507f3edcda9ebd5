// a7 / a8 / a9 / a10: the packed layout and what is computed through it.
//
//   rc.BinCount(pack=True) / rc.GroupFromBinCount / rc.GroupByPack32  (riptable/rt_numpy.py:1945-1972)
//       stable bucket sort of the row ids by bin = LSD radix sort on the low bits of iKey (radix.cu)
//   rc.GroupByAllPack32 first / nth / last                            (riptable/rt_numpy.py:1891)
//   rc.EmaAll32 GB_CUMSUM                                             (riptable/rt_grouping.py:3430)
//       segmented inclusive scan over the packed order, scattered back to row order
//   rc.MakeiFirst / rc.MakeiNext                                      (riptable/rt_numpy.py:1791-1854)
#include "common.cuh"
#include "cumsum_blocked.cuh"
#include "radix.cuh"
#include "values.cuh"

extern "C" int rtb_bincount(const void* ikey, int kdt, int64_t nrows, int64_t nbins, int32_t* out, void* stream);

namespace rtb {

// ---- iKey -> uint32 sort keys; rows outside [0, nbins) are counted (the caller rejects the input) ------------
__global__ void __launch_bounds__(256) pack_keys_kernel(const void* __restrict__ ikey, int kdt, int64_t n, int64_t nbins,
                                                        uint32_t* __restrict__ keys, uint32_t* bad) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  uint32_t nbad = 0;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; r < n; r += stride) {
    if (r + 3 < n) {
      int64_t k[4];
      load_ikey4(ikey, kdt, r, k);
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (k[i] < 0 || k[i] >= nbins) { nbad++; k[i] = 0; }
      *(uint4*)(keys + r) = make_uint4((uint32_t)k[0], (uint32_t)k[1], (uint32_t)k[2], (uint32_t)k[3]);
    } else {
      for (int64_t q = r; q < n; q++) {
        int64_t k = load_ikey(ikey, kdt, q);
        if (k < 0 || k >= nbins) { nbad++; k = 0; }
        keys[q] = (uint32_t)k;
      }
    }
  }
  if (nbad) atomicAdd(bad, nbad);
}

struct PackLayout {
  uint32_t *keys_a, *keys_b, *vals_a, *vals_b, *bad, *scan_scratch;
  void* radix_ws;
  size_t radix_bytes, bytes;
};
static PackLayout pack_layout(void* ws, size_t wsbytes, int64_t n, int64_t nbins) {
  Arena a(ws, wsbytes);
  PackLayout L;
  L.scan_scratch = a.take<uint32_t>(scan_u32_workspace_elems(nbins > 0 ? nbins : 1));
  size_t m = (size_t)(n > 0 ? n : 1);
  L.keys_a = a.take<uint32_t>(m);
  L.keys_b = a.take<uint32_t>(m);
  L.vals_a = a.take<uint32_t>(m);
  L.vals_b = a.take<uint32_t>(m);
  L.bad = a.take<uint32_t>(4);
  L.radix_bytes = radix_workspace_bytes(n);
  L.radix_ws = a.take<uint8_t>(L.radix_bytes);
  L.bytes = a.off;
  return L;
}

static int bits_for(int64_t nbins) {
  int b = 1;
  while (b < 32 && (1ll << b) < nbins) b++;
  return b;
}

// Sorts row ids by bin (stable).  On return *sorted_keys / *igroup point at the buffers (inside the workspace)
// holding the bins in packed order and the packed row ids.  Synchronises the stream.
static int pack_sort(const void* ikey, int kdt, int64_t n, int64_t nbins, const PackLayout& L, cudaStream_t stream,
                     uint32_t** sorted_keys, uint32_t** igroup, const char* who) {
  // pack_keys_kernel reads iKey four rows at a time: a misaligned column would fault and poison the context
  RTB_ARG(((uintptr_t)ikey % vec4_align(dtype_size(kdt))) == 0, "%s: ikey must be aligned to 4 rows (at most 16 bytes)", who);
  RTB_CUDA(cudaMemsetAsync(L.bad, 0, sizeof(uint32_t), stream));
  pack_keys_kernel<<<grid_for(n, 256, 4), 256, 0, stream>>>(ikey, kdt, n, nbins, L.keys_a, L.bad);
  RTB_LAUNCH_CHECK();
  uint32_t hbad = 0;
  RTB_CUDA(cudaMemcpyAsync(&hbad, L.bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  int in_b = 0;
  int rc = radix_sort_pairs<uint32_t>(L.keys_a, L.vals_a, L.keys_b, L.vals_b, n, bits_for(nbins), true, L.radix_ws,
                                      L.radix_bytes, stream, &in_b);
  if (rc != RTB_OK) return rc;
  RTB_CUDA(cudaStreamSynchronize(stream));
  RTB_ARG(hbad == 0, "%s: %u ikey values lie outside [0, %lld)", who, hbad, (long long)nbins);
  *sorted_keys = in_b ? L.keys_b : L.keys_a;
  *igroup = in_b ? L.vals_b : L.vals_a;
  return RTB_OK;
}

// run starts / ends of the sorted bins: start[k] = first packed position of bin k, end[k] = one past its last
__global__ void __launch_bounds__(256) run_bounds_kernel(const uint32_t* __restrict__ skeys, int64_t n, uint32_t* __restrict__ start,
                                                         uint32_t* __restrict__ end) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = skeys[j];
    if (j == 0 || skeys[j - 1] != k) start[k] = (uint32_t)j;
    if (j == n - 1 || skeys[j + 1] != k) end[k] = (uint32_t)(j + 1);
  }
}
__global__ void __launch_bounds__(256) run_lengths_kernel(const uint32_t* __restrict__ start, uint32_t* __restrict__ end_then_count, int64_t nbins) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbins; b += (int64_t)gridDim.x * blockDim.x)
    end_then_count[b] -= start[b];  // absent bins: 0 - 0
}

// item copy in the widest unit the item size allows (items start at multiples of their size inside 16-byte aligned buffers)
__device__ __forceinline__ void copy_item(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int itemsize) {
  if ((itemsize & 7) == 0) {
    for (int i = 0; i < itemsize; i += 8) *(unsigned long long*)(dst + i) = *(const unsigned long long*)(src + i);
  } else if ((itemsize & 3) == 0) {
    for (int i = 0; i < itemsize; i += 4) *(uint32_t*)(dst + i) = *(const uint32_t*)(src + i);
  } else if ((itemsize & 1) == 0) {
    for (int i = 0; i < itemsize; i += 2) *(uint16_t*)(dst + i) = *(const uint16_t*)(src + i);
  } else {
    for (int i = 0; i < itemsize; i++) dst[i] = src[i];
  }
}

// ---- a8: first / nth / last ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pick_kernel(const uint8_t* __restrict__ values, int vdt, int itemsize,
                                                   const int32_t* __restrict__ igroup, const int32_t* __restrict__ ifirst,
                                                   const int32_t* __restrict__ ncount, int64_t nb, int func, int64_t nth,
                                                   int64_t lo, int64_t hi, uint8_t* __restrict__ out) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    int64_t cnt = ncount[b];
    int64_t off = func == RTB_GB_FIRST ? 0 : func == RTB_GB_LAST ? cnt - 1 : (nth >= 0 ? nth : cnt + nth);
    bool ok = off >= 0 && off < cnt && b >= lo && b < hi;
    uint8_t* dst = out + b * (int64_t)itemsize;
    if (ok) {
      copy_item(dst, values + (int64_t)igroup[(int64_t)ifirst[b] + off] * itemsize, itemsize);
    } else if (vdt <= RTB_F64) {
      unsigned long long inv = invalid_raw(vdt);
      for (int i = 0; i < itemsize; i++) dst[i] = (uint8_t)(inv >> (8 * i));
    } else {
      for (int i = 0; i < itemsize; i++) dst[i] = 0;  // b'' / ''
    }
  }
}

__device__ __forceinline__ unsigned long long f64_bits(double d) { return (unsigned long long)__double_as_longlong(d); }
__device__ __forceinline__ double bits_f64(unsigned long long u) { return __longlong_as_double((long long)u); }

// ---- 8f.2: rolling ops on the packed layout (GroupByAllPack32 functions 200-206) ---------------------------
// One thread per packed position j: its row, its bin, and the window [max(start, j - w + 1), j] of the bin's run of
// iGroup.  Sums run oldest -> newest in f64 / int64, which is also the oracle's order (bit-exact).
struct RollArgs {
  const uint8_t* values;
  int vdt, itemsize, odt;
  const void* ikey;
  int kdt;
  const int32_t *igroup, *ifirst, *ncount;
  int64_t n, nb;
  int func;
  int64_t window;
  uint8_t* out;
};

__device__ __forceinline__ void store_invalid_item(uint8_t* dst, int dt, int itemsize) {
  if (dt <= RTB_F64) {
    const unsigned long long inv = invalid_raw(dt);
    switch (itemsize) {
      case 8: *(unsigned long long*)dst = inv; break;
      case 4: *(uint32_t*)dst = (uint32_t)inv; break;
      case 2: *(uint16_t*)dst = (uint16_t)inv; break;
      default: dst[0] = (uint8_t)inv;
    }
  } else {
    for (int i = 0; i < itemsize; i++) dst[i] = 0;
  }
}

__global__ void __launch_bounds__(256) rolling_kernel(RollArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int vsz = a.vdt <= RTB_F64 ? dsize(a.vdt) : a.itemsize;
  const bool flt = a.vdt == RTB_F32 || a.vdt == RTB_F64;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += stride) {
    const int64_t row = a.igroup[j];
    int64_t bin = load_ikey(a.ikey, a.kdt, row);
    if (bin < 0 || bin >= a.nb) bin = 0;
    const int64_t start = a.ifirst[bin], cnt = a.ncount[bin];
    switch (a.func) {
      case RTB_GB_ROLLING_COUNT: {
        const int64_t p = j - start;
        ((int32_t*)a.out)[row] = bin == 0 ? INT32_MIN : (int32_t)(a.window >= 0 ? p : cnt - 1 - p);
        break;
      }
      case RTB_GB_ROLLING_SHIFT: {
        const int64_t q = j - a.window;
        uint8_t* dst = a.out + row * (int64_t)a.itemsize;
        if (q >= start && q < start + cnt) {
          copy_item(dst, a.values + (int64_t)a.igroup[q] * a.itemsize, a.itemsize);
        } else {
          store_invalid_item(dst, a.vdt, a.itemsize);
        }
        break;
      }
      case RTB_GB_ROLLING_DIFF: {
        const int64_t q = j - a.window;
        unsigned long long raw = invalid_raw(a.odt);
        if (q >= start && q < start + cnt) {
          const unsigned long long x = load_raw1(a.values, vsz, row), y = load_raw1(a.values, vsz, a.igroup[q]);
          if (flt) {
            const double d = raw_to_f64(x, a.vdt) - raw_to_f64(y, a.vdt);
            raw = a.odt == RTB_F32 ? (unsigned long long)__float_as_uint((float)d) : f64_bits(d);
          } else {
            raw = (unsigned long long)(raw_to_i64(x, a.vdt) - raw_to_i64(y, a.vdt));
          }
        }
        store_raw(a.out, a.odt, row, raw);
        break;
      }
      default: {  // ROLLING_SUM / NANSUM / MEAN / NANMEAN
        const bool skip = a.func == RTB_GB_ROLLING_NANSUM || a.func == RTB_GB_ROLLING_NANMEAN;
        const bool mean = a.func == RTB_GB_ROLLING_MEAN || a.func == RTB_GB_ROLLING_NANMEAN;
        int64_t q = j - a.window + 1;
        if (q < start) q = start;
        double fs = 0.0;
        long long is = 0;
        int64_t used = 0;
        for (; q <= j; q++) {
          const unsigned long long x = load_raw1(a.values, vsz, a.igroup[q]);
          if (skip && raw_invalid(x, a.vdt)) continue;
          used++;
          if (flt || mean) fs += raw_to_f64(x, a.vdt);
          else is += raw_to_i64(x, a.vdt);
        }
        unsigned long long raw;
        if (mean) raw = f64_bits(used ? fs / (double)used : __longlong_as_double(0x7FF8000000000000ll));
        else if (a.odt == RTB_F32) raw = (unsigned long long)__float_as_uint((float)fs);
        else if (flt) raw = f64_bits(fs);
        else raw = (unsigned long long)is;
        store_raw(a.out, a.odt, row, raw);
      }
    }
  }
}

// ---- 8f.2: rolling nan-quantile (GroupByAllPack32 function 207) ---------------------------------------------
// One thread per packed position: the valid values of its window go to a local buffer, are sorted (insertion sort up
// to 32 values, Shell sort above) and the two order statistics around (m - 1) * q are averaged — numpy's "midpoint"
// rule in the (lower + higher) / 2 form riptable's own check uses (riptable/rt_numpy.py:4202-4254).
constexpr int kMaxQuantileWindow = 1024;

__global__ void __launch_bounds__(128) rolling_quantile_kernel(RollArgs a, double q) {
  double buf[kMaxQuantileWindow];
  const int vsz = dsize(a.vdt);
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = a.igroup[j];
    int64_t bin = load_ikey(a.ikey, a.kdt, row);
    if (bin < 0 || bin >= a.nb) bin = 0;
    const int64_t start = a.ifirst[bin];
    int64_t p = j - a.window + 1;
    if (p < start) p = start;
    int m = 0;
    for (; p <= j; p++) {
      const unsigned long long x = load_raw1(a.values, vsz, a.igroup[p]);
      if (!raw_invalid(x, a.vdt)) buf[m++] = raw_to_f64(x, a.vdt);
    }
    double r = nan;
    if (m > 0) {
      if (m <= 32) {
        for (int i = 1; i < m; i++) {
          const double x = buf[i];
          int k = i - 1;
          while (k >= 0 && buf[k] > x) { buf[k + 1] = buf[k]; k--; }
          buf[k + 1] = x;
        }
      } else {
        const int gaps[8] = {701, 301, 132, 57, 23, 10, 4, 1};
        for (int gi = 0; gi < 8; gi++) {
          const int gap = gaps[gi];
          for (int i = gap; i < m; i++) {
            const double x = buf[i];
            int k = i;
            while (k >= gap && buf[k - gap] > x) { buf[k] = buf[k - gap]; k -= gap; }
            buf[k] = x;
          }
        }
      }
      const double idx = (double)(m - 1) * q;
      const int il = (int)floor(idx), ih = (int)ceil(idx);
      r = (buf[il] + buf[ih]) / 2;
    }
    ((double*)a.out)[row] = r;
  }
}

// ---- a9: cumulative ops (cumsum / cumprod / cummin / cummax) -----------------------------------------------
constexpr int kSegThreads = 256;
constexpr int kSegItems = 8;
constexpr int kSegTile = kSegThreads * kSegItems;

// Scan operators.  Sums and products run on int64 / f64 payloads with plain identities; min / max run on the
// order-preserving uint64 codes of values.cuh and carry two state bits next to the segment-head flag, because
// "nothing seen yet" and "a NaN / sentinel was seen" (skipna=False) both have to come out as the invalid value.
enum { CUM_SUM_I = 0, CUM_SUM_F = 1, CUM_PROD_I = 2, CUM_PROD_F = 3, CUM_MIN = 4, CUM_MAX = 5 };
constexpr uint32_t SEG_HEAD = 1u, SEG_HAS = 2u, SEG_POISON = 4u;

struct SegPair {
  unsigned long long v;  // f64 bits, 64-bit integer, or value code
  uint32_t f;            // SEG_HEAD: a segment head lies inside the summarised range; SEG_HAS / SEG_POISON: min/max state
};


template <int OP>
__device__ __forceinline__ unsigned long long seg_op(unsigned long long a, unsigned long long b) {
  if (OP == CUM_SUM_F) return f64_bits(bits_f64(a) + bits_f64(b));
  if (OP == CUM_PROD_F) return f64_bits(bits_f64(a) * bits_f64(b));
  if (OP == CUM_PROD_I) return a * b;
  if (OP == CUM_MIN) return a < b ? a : b;
  if (OP == CUM_MAX) return a > b ? a : b;
  return a + b;
}
template <int OP>
__device__ __forceinline__ SegPair seg_identity() {
  SegPair r;
  r.f = 0;
  r.v = OP == CUM_PROD_I ? 1ull : OP == CUM_PROD_F ? 0x3FF0000000000000ull : 0ull;
  return r;
}
template <int OP>
__device__ __forceinline__ SegPair seg_combine(SegPair a, SegPair b) {  // a precedes b
  SegPair r;
  if (OP < CUM_MIN) {
    r.f = a.f | b.f;
    r.v = b.f ? b.v : seg_op<OP>(a.v, b.v);
  } else {
    if (b.f & SEG_HEAD) return b;
    r.f = a.f | b.f;
    r.v = (a.f & b.f & SEG_HAS) ? seg_op<OP>(a.v, b.v) : (a.f & SEG_HAS) ? a.v : b.v;
  }
  return r;
}

// inclusive scan of SegPair across the block; returns this thread's EXCLUSIVE prefix and the block total
template <int OP>
__device__ __forceinline__ SegPair seg_block_exclusive(SegPair x, SegPair* total) {
  __shared__ SegPair wtot[kSegThreads / 32];
  __shared__ SegPair btot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  SegPair inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    SegPair t;
    t.v = __shfl_up_sync(0xffffffffu, inc.v, d);
    t.f = __shfl_up_sync(0xffffffffu, inc.f, d);
    if (lane >= d) inc = seg_combine<OP>(t, inc);
  }
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  SegPair ex = seg_identity<OP>();  // exclusive prefix = (previous warps) o (previous lanes)
  for (int i = 0; i < w; i++) ex = seg_combine<OP>(ex, wtot[i]);
  SegPair prev;
  prev.v = __shfl_up_sync(0xffffffffu, inc.v, 1);
  prev.f = __shfl_up_sync(0xffffffffu, inc.f, 1);
  if (lane > 0) ex = seg_combine<OP>(ex, prev);
  if (threadIdx.x == kSegThreads - 1) btot = seg_combine<OP>(ex, x);
  __syncthreads();
  *total = btot;
  __syncthreads();
  return ex;
}

struct CumArgs {
  const void* values;
  int vdt, odt;
  int skipna;              // min/max: invalid values are skipped (1) or poison the rest of the segment (0)
  const uint32_t* skeys;   // bins in packed order
  const uint32_t* igroup;  // packed row ids
  const uint8_t* filter;
  const uint8_t* reset;
  int64_t n;
  unsigned long long* xs;  // gathered operands in packed order
  uint8_t* heads;          // SEG_* flags in packed order
  SegPair* tile_agg;
  void* out;
};

// pass 1: gather operands, mark heads, reduce each tile
template <int OP>
__global__ void __launch_bounds__(kSegThreads) cumop_gather_kernel(CumArgs a) {
  const int64_t base = (int64_t)blockIdx.x * kSegTile + (int64_t)threadIdx.x * kSegItems;
  SegPair acc = seg_identity<OP>();
  const int vsz = dsize(a.vdt);
#pragma unroll
  for (int i = 0; i < kSegItems; i++) {
    int64_t j = base + i;
    if (j < a.n) {
      uint32_t row = a.igroup[j];
      uint32_t k = a.skeys[j];
      bool add = k != 0 && (!a.filter || a.filter[row]);
      bool head = j == 0 || a.skeys[j - 1] != k || (a.reset && add && a.reset[row]);
      SegPair p = seg_identity<OP>();
      p.f = head ? SEG_HEAD : 0u;
      if (add) {
        unsigned long long raw = load_raw1(a.values, vsz, row);
        if (OP == CUM_SUM_F || OP == CUM_PROD_F) p.v = f64_bits(raw_to_f64(raw, a.vdt));
        else if (OP == CUM_SUM_I || OP == CUM_PROD_I) p.v = (unsigned long long)raw_to_i64(raw, a.vdt);
        else if (raw_invalid(raw, a.vdt)) p.f |= a.skipna ? 0u : SEG_POISON;
        else { p.v = raw_to_code(raw, a.vdt, OP == CUM_MAX); p.f |= SEG_HAS; }
      }
      a.xs[j] = p.v;
      a.heads[j] = (uint8_t)p.f;
      acc = seg_combine<OP>(acc, p);
    }
  }
  SegPair total;
  seg_block_exclusive<OP>(acc, &total);
  if (threadIdx.x == 0) a.tile_agg[blockIdx.x] = total;
}

// pass 2 (one block): tile_agg[t] <- combination of all tiles before t
template <int OP>
__global__ void __launch_bounds__(kSegThreads) cumop_tiles_kernel(SegPair* tile_agg, int64_t ntiles) {
  SegPair carry = seg_identity<OP>();
  for (int64_t b = 0; b < ntiles; b += kSegThreads) {
    int64_t t = b + threadIdx.x;
    SegPair x = seg_identity<OP>();
    if (t < ntiles) x = tile_agg[t];
    SegPair total;
    SegPair ex = seg_block_exclusive<OP>(x, &total);
    if (t < ntiles) tile_agg[t] = seg_combine<OP>(carry, ex);
    carry = seg_combine<OP>(carry, total);
  }
}

// pass 3: running values inside each tile, scattered to row order
template <int OP>
__global__ void __launch_bounds__(kSegThreads) cumop_apply_kernel(CumArgs a) {
  const int64_t base = (int64_t)blockIdx.x * kSegTile + (int64_t)threadIdx.x * kSegItems;
  SegPair p[kSegItems];
  SegPair acc = seg_identity<OP>();
#pragma unroll
  for (int i = 0; i < kSegItems; i++) {
    int64_t j = base + i;
    p[i] = seg_identity<OP>();
    if (j < a.n) {
      p[i].v = a.xs[j];
      p[i].f = a.heads[j];
      acc = seg_combine<OP>(acc, p[i]);
    }
  }
  SegPair total;
  SegPair run = seg_combine<OP>(a.tile_agg[blockIdx.x], seg_block_exclusive<OP>(acc, &total));
#pragma unroll
  for (int i = 0; i < kSegItems; i++) {
    int64_t j = base + i;
    if (j < a.n) {
      run = seg_combine<OP>(run, p[i]);
      uint32_t row = a.igroup[j];
      unsigned long long raw;
      if (a.skeys[j] == 0) raw = invalid_raw(a.odt);
      else if (OP >= CUM_MIN) {
        raw = (run.f & (SEG_HAS | SEG_POISON)) == SEG_HAS ? code_to_raw(run.v, a.vdt, OP == CUM_MAX) : invalid_raw(a.odt);
      } else if (a.odt == RTB_F32) raw = __float_as_uint((float)bits_f64(run.v));
      else raw = run.v;
      store_raw(a.out, a.odt, row, raw);
    }
  }
}

template <int OP>
static int cumop_launch(const CumArgs& a, int64_t ntiles, cudaStream_t stream) {
  cumop_gather_kernel<OP><<<(unsigned)ntiles, kSegThreads, 0, stream>>>(a);
  RTB_LAUNCH_CHECK();
  cumop_tiles_kernel<OP><<<1, kSegThreads, 0, stream>>>(a.tile_agg, ntiles);
  RTB_LAUNCH_CHECK();
  cumop_apply_kernel<OP><<<(unsigned)ntiles, kSegThreads, 0, stream>>>(a);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

// ---- 8f.2: time-decayed moving averages (EmaAll32 functions 301 / 304 / 305) ---------------------------------
// Every recurrence is y_i = a_i * y_{i-1} + b_i over the packed order, so one inclusive scan of the affine maps under
// composition (a1, b1) o (a2, b2) = (a1 a2, a2 b1 + b2) does all groups at once.  A row that restarts its group's state
// (first row of a bin, or reset_filter) is a HEAD: it must cut off everything before it EXACTLY -- "a = 0" alone does
// not, because 0 * NaN and 0 * Inf are NaN and one NaN / Inf value (or an overflowing exp(-rate * dt) on unsorted
// times) would leak into every later bin.  A head is therefore marked by a = -0.0 (no decay weight is ever negative
// zero), `combine` returns a head on the right-hand side untouched, and -0.0 * w keeps the mark on compositions that
// start with one.  The identity (1, 0) is likewise passed through exactly (w = Inf would turn its 0 into NaN).  Rows
// the op-filter excludes are the identity
// (1, 0): they show the group's current value and leave its clock alone, so the decay of the next included row spans
// the time since the previous INCLUDED row -- found with a plain max-scan over (included ? position : -1).
struct Affine {
  double a, b;
};
struct AffineOp {
  typedef Affine T;
  __device__ static Affine identity() { return Affine{1.0, 0.0}; }
  __device__ static bool is_head(Affine m) { return __double_as_longlong(m.a) == (long long)0x8000000000000000ull; }
  __device__ static Affine head(double b) { return Affine{-0.0, b}; }
  __device__ static Affine combine(Affine x, Affine y) {  // x first, then y
    if (is_head(y) || (x.a == 1.0 && x.b == 0.0)) return y;
    return Affine{x.a * y.a, y.a * x.b + y.b};
  }
  __device__ static Affine shfl_up(Affine v, int d) { return Affine{__shfl_up_sync(0xffffffffu, v.a, d), __shfl_up_sync(0xffffffffu, v.b, d)}; }
};
struct MaxI32Op {
  typedef int32_t T;
  __device__ static int32_t identity() { return -1; }
  __device__ static int32_t combine(int32_t x, int32_t y) { return x > y ? x : y; }
  __device__ static int32_t shfl_up(int32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
};

// exclusive prefix of this thread's value across the block (block total in *total)
template <typename Op>
__device__ __forceinline__ typename Op::T plain_block_exclusive(typename Op::T x, typename Op::T* total) {
  typedef typename Op::T T;
  __shared__ T wtot[kSegThreads / 32];
  __shared__ T btot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = Op::shfl_up(inc, d);
    if (lane >= d) inc = Op::combine(t, inc);
  }
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  T ex = Op::identity();
  for (int i = 0; i < w; i++) ex = Op::combine(ex, wtot[i]);
  T prev = Op::shfl_up(inc, 1);
  if (lane > 0) ex = Op::combine(ex, prev);
  if (threadIdx.x == kSegThreads - 1) btot = Op::combine(ex, x);
  __syncthreads();
  *total = btot;
  __syncthreads();
  return ex;
}

// Plain inclusive scan, in place, three launches: tile totals, scan of the totals (one block), apply.
template <typename Op>
__global__ void __launch_bounds__(kSegThreads) plain_tiles_kernel(const typename Op::T* __restrict__ x, int64_t n, typename Op::T* tile_agg) {
  typedef typename Op::T T;
  const int64_t base = (int64_t)blockIdx.x * kSegTile + (int64_t)threadIdx.x * kSegItems;
  T acc = Op::identity();
#pragma unroll
  for (int i = 0; i < kSegItems; i++)
    if (base + i < n) acc = Op::combine(acc, x[base + i]);
  T total;
  plain_block_exclusive<Op>(acc, &total);
  if (threadIdx.x == 0) tile_agg[blockIdx.x] = total;
}
template <typename Op>
__global__ void __launch_bounds__(kSegThreads) plain_scan_tiles_kernel(typename Op::T* tile_agg, int64_t ntiles) {
  typedef typename Op::T T;
  T carry = Op::identity();
  for (int64_t b = 0; b < ntiles; b += kSegThreads) {
    const int64_t t = b + threadIdx.x;
    T x = t < ntiles ? tile_agg[t] : Op::identity();
    T total;
    T ex = plain_block_exclusive<Op>(x, &total);
    if (t < ntiles) tile_agg[t] = Op::combine(carry, ex);
    carry = Op::combine(carry, total);
  }
}
template <typename Op>
__global__ void __launch_bounds__(kSegThreads) plain_apply_kernel(typename Op::T* __restrict__ x, int64_t n, const typename Op::T* __restrict__ tile_agg) {
  typedef typename Op::T T;
  const int64_t base = (int64_t)blockIdx.x * kSegTile + (int64_t)threadIdx.x * kSegItems;
  T v[kSegItems];
  T acc = Op::identity();
#pragma unroll
  for (int i = 0; i < kSegItems; i++) {
    v[i] = base + i < n ? x[base + i] : Op::identity();
    acc = Op::combine(acc, v[i]);
  }
  T total;
  T run = Op::combine(tile_agg[blockIdx.x], plain_block_exclusive<Op>(acc, &total));
#pragma unroll
  for (int i = 0; i < kSegItems; i++)
    if (base + i < n) {
      run = Op::combine(run, v[i]);
      x[base + i] = run;
    }
}
template <typename Op>
static int plain_inclusive_scan(typename Op::T* x, int64_t n, typename Op::T* tile_agg, cudaStream_t stream) {
  const int64_t ntiles = (n + kSegTile - 1) / kSegTile;
  plain_tiles_kernel<Op><<<(unsigned)ntiles, kSegThreads, 0, stream>>>(x, n, tile_agg);
  RTB_LAUNCH_CHECK();
  plain_scan_tiles_kernel<Op><<<1, kSegThreads, 0, stream>>>(tile_agg, ntiles);
  RTB_LAUNCH_CHECK();
  plain_apply_kernel<Op><<<(unsigned)ntiles, kSegThreads, 0, stream>>>(x, n, tile_agg);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

struct EmaArgs {
  const void* values;
  int vdt, odt;
  const void* time;
  int tdt;
  int func;
  double rate;
  const uint32_t* skeys;
  const uint32_t* igroup;
  const uint8_t* filter;
  const uint8_t* reset;
  int64_t n;
  int32_t* lastinc;  // in: included ? position : -1; after the max-scan: last included position <= j
  Affine* maps;
  void* out;
};

__global__ void __launch_bounds__(256) ema_mark_kernel(EmaArgs a) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * blockDim.x) {
    const bool inc = a.skeys[j] != 0 && (!a.filter || a.filter[a.igroup[j]]);
    a.lastinc[j] = inc ? (int32_t)j : -1;
  }
}

__global__ void __launch_bounds__(256) ema_maps_kernel(EmaArgs a) {
  const int vsz = dsize(a.vdt), tsz = a.time ? dsize(a.tdt) : 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = a.skeys[j];
    const bool seg_first = j == 0 || a.skeys[j - 1] != k;
    Affine m;
    if (a.lastinc[j] != (int32_t)j) {  // bin 0, or excluded by the op-filter: show the running value
      m = seg_first ? AffineOp::head(0.0) : Affine{1.0, 0.0};
    } else {
      const uint32_t row = a.igroup[j];
      const double x = raw_to_f64(load_raw1(a.values, vsz, row), a.vdt);
      const int32_t prev = j > 0 ? a.lastinc[j - 1] : -1;
      const bool restart = prev < 0 || a.skeys[prev] != k || (a.reset && a.reset[row]);
      if (restart) {
        m = AffineOp::head(x);
      } else if (a.func == RTB_GB_EMAWEIGHTED) {
        m = Affine{a.rate, x * (1.0 - a.rate)};
      } else {
        const double t = raw_to_f64(load_raw1(a.time, tsz, row), a.tdt);
        const double tp = raw_to_f64(load_raw1(a.time, tsz, a.igroup[prev]), a.tdt);
        const double w = exp(-a.rate * (t - tp));
        m = a.func == RTB_GB_EMADECAY ? Affine{w, x} : Affine{w, x * (1.0 - w)};
      }
    }
    a.maps[j] = m;
  }
}

__global__ void __launch_bounds__(256) ema_store_kernel(EmaArgs a) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t row = a.igroup[j];
    unsigned long long raw;
    if (a.skeys[j] == 0) raw = invalid_raw(a.odt);
    else if (a.odt == RTB_F32) raw = __float_as_uint((float)a.maps[j].b);
    else raw = f64_bits(a.maps[j].b);
    store_raw(a.out, a.odt, row, raw);
  }
}

struct EmaLayout {
  PackLayout pack;
  void* pack_ws;
  int32_t *lastinc, *tile_i;
  Affine *maps, *tile_a;
  size_t pack_bytes, bytes;
};
static EmaLayout ema_layout(void* ws, size_t wsbytes, int64_t n) {
  Arena a(ws, wsbytes);
  EmaLayout L;
  L.pack_bytes = pack_layout(nullptr, 0, n, 1).bytes + 256;
  L.pack_ws = a.take<uint8_t>(L.pack_bytes);
  L.pack = pack_layout(L.pack_ws, L.pack_bytes, n, 1);
  const size_t m = (size_t)(n > 0 ? n : 1), tiles = (m + kSegTile - 1) / kSegTile + 1;
  L.lastinc = a.take<int32_t>(m);
  L.maps = a.take<Affine>(m);
  L.tile_i = a.take<int32_t>(tiles);
  L.tile_a = a.take<Affine>(tiles);
  L.bytes = a.off;
  return L;
}

struct CumLayout {
  PackLayout pack;
  void* pack_ws;
  unsigned long long* xs;
  uint8_t* heads;
  SegPair* tile_agg;
  size_t pack_bytes, bytes;
};
static CumLayout cum_layout(void* ws, size_t wsbytes, int64_t n) {
  Arena a(ws, wsbytes);
  CumLayout L;
  L.pack_bytes = pack_layout(nullptr, 0, n, 1).bytes + 256;
  L.pack_ws = a.take<uint8_t>(L.pack_bytes);
  L.pack = pack_layout(L.pack_ws, L.pack_bytes, n, 1);
  size_t m = (size_t)(n > 0 ? n : 1);
  L.xs = a.take<unsigned long long>(m);
  L.heads = a.take<uint8_t>(m);
  L.tile_agg = a.take<SegPair>((m + kSegTile - 1) / kSegTile + 1);
  L.bytes = a.off;
  return L;
}

// ---- a10 ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ifirst_scan_kernel(const void* __restrict__ ikey, int kdt, int64_t n, int64_t nmax,
                                                          const uint8_t* __restrict__ filter, int mode, uint32_t* tmp) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    int64_t k = load_ikey(ikey, kdt, r);
    if (k > 0 && k <= nmax && (!filter || filter[r])) {
      if (mode == 0) atomicMin(&tmp[k], (uint32_t)r);
      else atomicMax(&tmp[k], (uint32_t)r + 1u);
    }
  }
}
__global__ void __launch_bounds__(256) ifirst_finish_kernel(uint32_t* tmp, int64_t nslots, int mode) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nslots; b += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = tmp[b];
    int32_t o;
    if (mode == 0) o = v == 0xFFFFFFFFu ? INT32_MIN : (int32_t)v;
    else o = v == 0u ? INT32_MIN : (int32_t)(v - 1u);
    if (b == 0) o = INT32_MIN;
    ((int32_t*)tmp)[b] = o;
  }
}

__global__ void __launch_bounds__(256) inext_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ igroup,
                                                    int64_t n, int mode, int32_t* __restrict__ out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    uint32_t k = skeys[j];
    int32_t o = INT32_MIN;
    if (k != 0) {
      if (mode == 0) {
        if (j + 1 < n && skeys[j + 1] == k) o = (int32_t)igroup[j + 1];
      } else {
        if (j > 0 && skeys[j - 1] == k) o = (int32_t)igroup[j - 1];
      }
    }
    out[igroup[j]] = o;
  }
}

}  // namespace rtb

using namespace rtb;

extern "C" size_t rtb_pack_workspace_bytes(int64_t nrows, int64_t nbins) {
  if (nrows < 0) return 0;
  return pack_layout(nullptr, 0, nrows, nbins).bytes + 256;
}

extern "C" int rtb_groupby_pack(const void* ikey, int kdt, int64_t nrows, int64_t nbins, int32_t* igroup_out,
                                int32_t* ifirstgroup_out, int32_t* ncountgroup_out, void* workspace, size_t workspace_bytes,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(is_ikey_dtype(kdt), "pack: ikey must be int8/16/32/64");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && nbins >= 1 && nbins < (1ll << 31), "pack: bad sizes");
  RTB_ARG(ifirstgroup_out && ncountgroup_out, "pack: null output");
  PackLayout L = pack_layout(workspace, workspace_bytes, nrows, nbins);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "pack: workspace too small");
  // nCountGroup from the run lengths of the sorted bins (two streaming passes) instead of one atomic per row
  RTB_CUDA(cudaMemsetAsync(ncountgroup_out, 0, sizeof(int32_t) * (size_t)nbins, stream));
  RTB_CUDA(cudaMemsetAsync(ifirstgroup_out, 0, sizeof(int32_t) * (size_t)nbins, stream));
  int rc = RTB_OK;
  if (nrows > 0) {
    RTB_ARG(ikey && igroup_out, "pack: null pointer");
    uint32_t *skeys, *ig;
    rc = pack_sort(ikey, kdt, nrows, nbins, L, stream, &skeys, &ig, "pack");
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaMemcpyAsync(igroup_out, ig, sizeof(int32_t) * (size_t)nrows, cudaMemcpyDeviceToDevice, stream));
    run_bounds_kernel<<<grid_for(nrows, 256, 4), 256, 0, stream>>>(skeys, nrows, (uint32_t*)ifirstgroup_out, (uint32_t*)ncountgroup_out);
    RTB_LAUNCH_CHECK();
    run_lengths_kernel<<<grid_for(nbins, 256, 4), 256, 0, stream>>>((const uint32_t*)ifirstgroup_out, (uint32_t*)ncountgroup_out, nbins);
    RTB_LAUNCH_CHECK();
  }
  // iFirstGroup = exclusive cumsum(nCountGroup)
  return exclusive_scan_u32((const uint32_t*)ncountgroup_out, (uint32_t*)ifirstgroup_out, nbins, SCAN_IDENTITY, L.scan_scratch, nullptr, stream);
}

extern "C" int rtb_groupby_pick(const void* values, int vdt, int32_t itemsize, const int32_t* igroup,
                                const int32_t* ifirstgroup, const int32_t* ncountgroup, int64_t unique_rows, int func,
                                int64_t nth, int64_t bin_low, int64_t bin_high, void* out, void* stream) {
  RTB_ARG(func == RTB_GB_FIRST || func == RTB_GB_NTH || func == RTB_GB_LAST, "GroupByAllPack32: function %d is not first/nth/last", func);
  RTB_ARG(unique_rows >= 0 && itemsize >= 1 && out && ifirstgroup && ncountgroup, "GroupByAllPack32: bad arguments");
  RTB_ARG(vdt >= RTB_BOOL && vdt <= RTB_UCS4, "GroupByAllPack32: bad dtype");
  RTB_ARG(vdt > RTB_F64 || itemsize == dtype_size(vdt), "GroupByAllPack32: itemsize does not match dtype");
  int64_t nb = unique_rows + 1;
  pick_kernel<<<grid_for(nb, 256, 1), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)values, vdt, itemsize, igroup, ifirstgroup,
                                                                      ncountgroup, nb, func, nth, bin_low, bin_high, (uint8_t*)out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_groupby_rolling_quantile(const void* values, int vdt, const void* ikey, int kdt, const int32_t* igroup,
                                            const int32_t* ifirstgroup, const int32_t* ncountgroup, int64_t nrows,
                                            int64_t unique_rows, int64_t window, double q, double* out, void* stream) {
  RTB_ARG(is_ikey_dtype(kdt), "GroupByAllPack32: ikey must be int8/16/32/64");
  if (vdt < RTB_BOOL || vdt > RTB_F64) {
    set_error("GroupByAllPack32: rolling_quantile is not provided for dtype %d", vdt);
    return RTB_ERR_UNSUPPORTED;
  }
  RTB_ARG(q >= 0.0 && q <= 1.0, "GroupByAllPack32: quantile must lie in [0, 1]");
  RTB_ARG(window >= 1, "GroupByAllPack32: rolling window must be >= 1");
  RTB_ARG(window <= kMaxQuantileWindow, "GroupByAllPack32: rolling_quantile windows above %d rows are not provided", kMaxQuantileWindow);
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && unique_rows >= 0, "GroupByAllPack32: bad sizes");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(values && ikey && igroup && ifirstgroup && ncountgroup && out, "GroupByAllPack32: null pointer");
  RollArgs a;
  a.values = (const uint8_t*)values; a.vdt = vdt; a.itemsize = dtype_size(vdt); a.odt = RTB_F64; a.ikey = ikey; a.kdt = kdt;
  a.igroup = igroup; a.ifirst = ifirstgroup; a.ncount = ncountgroup; a.n = nrows; a.nb = unique_rows + 1;
  a.func = RTB_GB_ROLLING_QUANTILE; a.window = window; a.out = (uint8_t*)out;
  rolling_quantile_kernel<<<grid_for(nrows, 128, 1), 128, 0, (cudaStream_t)stream>>>(a, q);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_cumop_out_dtype(int func, int vdt);
extern "C" int rtb_rolling_out_dtype(int func, int vdt) {
  switch (func) {
    case RTB_GB_ROLLING_COUNT: return RTB_I32;
    case RTB_GB_ROLLING_SHIFT: return (vdt >= RTB_BOOL && vdt <= RTB_UCS4) ? vdt : RTB_ERR_UNSUPPORTED;
    case RTB_GB_ROLLING_DIFF: return (vdt > RTB_BOOL && vdt <= RTB_F64) ? vdt : RTB_ERR_UNSUPPORTED;
    case RTB_GB_ROLLING_SUM: case RTB_GB_ROLLING_NANSUM: return rtb_cumop_out_dtype(RTB_GB_CUMSUM, vdt);
    case RTB_GB_ROLLING_MEAN: case RTB_GB_ROLLING_NANMEAN: return (vdt >= RTB_BOOL && vdt <= RTB_F64) ? (int)RTB_F64 : (int)RTB_ERR_UNSUPPORTED;
    default: return RTB_ERR_UNSUPPORTED;
  }
}

extern "C" int rtb_groupby_rolling(int func, const void* values, int vdt, int32_t itemsize, const void* ikey, int kdt,
                                   const int32_t* igroup, const int32_t* ifirstgroup, const int32_t* ncountgroup,
                                   int64_t nrows, int64_t unique_rows, int64_t window, void* out, void* stream) {
  RTB_ARG(is_ikey_dtype(kdt), "GroupByAllPack32: ikey must be int8/16/32/64");
  const int odt = rtb_rolling_out_dtype(func, func == RTB_GB_ROLLING_COUNT ? RTB_I32 : vdt);
  if (odt < 0) {
    set_error("GroupByAllPack32: rolling function %d is not provided for dtype %d", func, vdt);
    return RTB_ERR_UNSUPPORTED;
  }
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && unique_rows >= 0, "GroupByAllPack32: bad sizes");
  const bool windowed = func != RTB_GB_ROLLING_SHIFT && func != RTB_GB_ROLLING_DIFF && func != RTB_GB_ROLLING_COUNT;
  RTB_ARG(!windowed || window >= 1, "GroupByAllPack32: rolling window must be >= 1");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(ikey && igroup && ifirstgroup && ncountgroup && out, "GroupByAllPack32: null pointer");
  RTB_ARG(func == RTB_GB_ROLLING_COUNT || values, "GroupByAllPack32: null values");
  RTB_ARG(func == RTB_GB_ROLLING_COUNT || vdt > RTB_F64 || itemsize == dtype_size(vdt), "GroupByAllPack32: itemsize does not match dtype");
  RollArgs a;
  a.values = (const uint8_t*)values; a.vdt = vdt; a.itemsize = itemsize; a.odt = odt; a.ikey = ikey; a.kdt = kdt;
  a.igroup = igroup; a.ifirst = ifirstgroup; a.ncount = ncountgroup; a.n = nrows; a.nb = unique_rows + 1; a.func = func;
  a.window = window; a.out = (uint8_t*)out;
  rolling_kernel<<<grid_for(nrows, 256, 1), 256, 0, (cudaStream_t)stream>>>(a);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_cumop_out_dtype(int func, int vdt) {
  if (vdt < RTB_BOOL || vdt > RTB_F64) return RTB_ERR_UNSUPPORTED;
  switch (func) {
    case RTB_GB_CUMSUM: case RTB_GB_CUMPROD:
      if (vdt == RTB_F32 || vdt == RTB_F64 || vdt == RTB_U64) return vdt;
      return RTB_I64;
    case RTB_GB_CUMMIN: case RTB_GB_CUMMAX: case RTB_GB_CUMNANMIN: case RTB_GB_CUMNANMAX: return vdt;
    default: return RTB_ERR_UNSUPPORTED;
  }
}
extern "C" int rtb_cumsum_out_dtype(int vdt) { return rtb_cumop_out_dtype(RTB_GB_CUMSUM, vdt); }

extern "C" size_t rtb_cumsum_workspace_bytes(int64_t nrows, int64_t unique_rows) {
  if (nrows < 0) return 0;
  const size_t sorted = cum_layout(nullptr, 0, nrows).bytes + 256;
  const CbPlan bp = cumsum_blocked_plan(RTB_GB_CUMSUM, nrows, unique_rows, false);
  return bp.ok && bp.bytes + 256 > sorted ? bp.bytes + 256 : sorted;  // enough for either form
}

// What one call needs: the blocked form's chunk matrix when it applies (GB_CUMSUM, no reset filter, up to 4 Mi bins),
// the packed layout otherwise.
extern "C" size_t rtb_cumop_workspace_bytes(int func, int64_t nrows, int64_t unique_rows, int has_reset_filter) {
  if (nrows < 0) return 0;
  const CbPlan bp = cumsum_blocked_plan(func, nrows, unique_rows, has_reset_filter != 0);
  if (bp.ok) return bp.bytes + 256;
  return cum_layout(nullptr, 0, nrows).bytes + 256;
}

extern "C" int rtb_groupby_cumop(int func, const void* values, int vdt, const void* ikey, int kdt, int64_t nrows,
                                 int64_t unique_rows, const uint8_t* filter, const uint8_t* reset_filter, void* out,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(is_ikey_dtype(kdt), "EmaAll32: ikey must be int8/16/32/64");
  int odt = rtb_cumop_out_dtype(func, vdt);
  if (odt < 0) {
    set_error("EmaAll32: function %d is not provided for dtype %d", func, vdt);
    return RTB_ERR_UNSUPPORTED;
  }
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && unique_rows >= 0, "EmaAll32: bad sizes");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(values && ikey && out, "EmaAll32: null pointer");
  RTB_ARG(((uintptr_t)ikey % vec4_align(dtype_size(kdt))) == 0, "EmaAll32: ikey must be aligned to 4 rows");
  const CbPlan bp = cumsum_blocked_plan(func, nrows, unique_rows, reset_filter != nullptr);
  if (bp.ok && workspace && bp.bytes <= workspace_bytes && ((uintptr_t)workspace % 16) == 0) {
    prof_begin(PROF_SEGSCAN, stream);
    int brc = cumsum_blocked_run(bp, func, values, vdt, odt, ikey, kdt, nrows, unique_rows, filter, out, workspace, stream);
    if (brc != RTB_OK) return brc;
    prof_end(PROF_SEGSCAN, stream, (uint64_t)nrows);
    return RTB_OK;
  }
  CumLayout L = cum_layout(workspace, workspace_bytes, nrows);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "EmaAll32: workspace too small");
  uint32_t *skeys, *ig;
  int rc = pack_sort(ikey, kdt, nrows, unique_rows + 1, L.pack, stream, &skeys, &ig, "EmaAll32");
  if (rc != RTB_OK) return rc;
  CumArgs a;
  a.values = values; a.vdt = vdt; a.odt = odt; a.skeys = skeys; a.igroup = ig; a.filter = filter; a.reset = reset_filter;
  a.skipna = func == RTB_GB_CUMNANMIN || func == RTB_GB_CUMNANMAX;
  a.n = nrows; a.xs = L.xs; a.heads = L.heads; a.tile_agg = L.tile_agg; a.out = out;
  const int64_t ntiles = (nrows + kSegTile - 1) / kSegTile;
  const bool fl = vdt == RTB_F32 || vdt == RTB_F64;
  prof_begin(PROF_SEGSCAN, stream);
  switch (func) {
    case RTB_GB_CUMSUM: rc = fl ? cumop_launch<CUM_SUM_F>(a, ntiles, stream) : cumop_launch<CUM_SUM_I>(a, ntiles, stream); break;
    case RTB_GB_CUMPROD: rc = fl ? cumop_launch<CUM_PROD_F>(a, ntiles, stream) : cumop_launch<CUM_PROD_I>(a, ntiles, stream); break;
    case RTB_GB_CUMMIN: case RTB_GB_CUMNANMIN: rc = cumop_launch<CUM_MIN>(a, ntiles, stream); break;
    default: rc = cumop_launch<CUM_MAX>(a, ntiles, stream); break;
  }
  if (rc != RTB_OK) return rc;
  prof_end(PROF_SEGSCAN, stream, (uint64_t)nrows);
  return RTB_OK;
}

extern "C" int rtb_groupby_cumsum(const void* values, int vdt, const void* ikey, int kdt, int64_t nrows, int64_t unique_rows,
                                  const uint8_t* filter, const uint8_t* reset_filter, void* out, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  return rtb_groupby_cumop(RTB_GB_CUMSUM, values, vdt, ikey, kdt, nrows, unique_rows, filter, reset_filter, out, workspace,
                           workspace_bytes, stream_);
}

extern "C" int rtb_ema_out_dtype(int func, int vdt) {
  if (func != RTB_GB_EMADECAY && func != RTB_GB_EMANORMAL && func != RTB_GB_EMAWEIGHTED) return RTB_ERR_UNSUPPORTED;
  if (vdt < RTB_BOOL || vdt > RTB_F64) return RTB_ERR_UNSUPPORTED;
  return vdt == RTB_F32 ? RTB_F32 : RTB_F64;
}

extern "C" size_t rtb_ema_workspace_bytes(int64_t nrows, int64_t unique_rows) {
  (void)unique_rows;
  if (nrows < 0) return 0;
  return ema_layout(nullptr, 0, nrows).bytes + 256;
}

extern "C" int rtb_groupby_ema(int func, const void* values, int vdt, const void* ikey, int kdt, int64_t nrows,
                               int64_t unique_rows, const void* time, int tdt, double decay_rate, const uint8_t* filter,
                               const uint8_t* reset_filter, void* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(is_ikey_dtype(kdt), "EmaAll32: ikey must be int8/16/32/64");
  const int odt = rtb_ema_out_dtype(func, vdt);
  if (odt < 0) {
    set_error("EmaAll32: function %d is not provided for dtype %d", func, vdt);
    return RTB_ERR_UNSUPPORTED;
  }
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && unique_rows >= 0, "EmaAll32: bad sizes");
  RTB_ARG(func == RTB_GB_EMAWEIGHTED || (time && tdt >= RTB_BOOL && tdt <= RTB_F64), "EmaAll32: ema_decay / ema_normal need a numeric time array");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(values && ikey && out, "EmaAll32: null pointer");
  RTB_ARG(((uintptr_t)ikey % vec4_align(dtype_size(kdt))) == 0, "EmaAll32: ikey must be aligned to 4 rows");
  EmaLayout L = ema_layout(workspace, workspace_bytes, nrows);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "EmaAll32: workspace too small");
  uint32_t *skeys, *ig;
  int rc = pack_sort(ikey, kdt, nrows, unique_rows + 1, L.pack, stream, &skeys, &ig, "EmaAll32");
  if (rc != RTB_OK) return rc;
  EmaArgs a;
  a.values = values; a.vdt = vdt; a.odt = odt; a.time = func == RTB_GB_EMAWEIGHTED ? nullptr : time; a.tdt = tdt; a.func = func;
  a.rate = decay_rate; a.skeys = skeys; a.igroup = ig; a.filter = filter; a.reset = reset_filter; a.n = nrows;
  a.lastinc = L.lastinc; a.maps = L.maps; a.out = out;
  const int g = grid_for(nrows, 256, 4);
  prof_begin(PROF_SEGSCAN, stream);
  ema_mark_kernel<<<g, 256, 0, stream>>>(a);
  RTB_LAUNCH_CHECK();
  rc = plain_inclusive_scan<MaxI32Op>(L.lastinc, nrows, L.tile_i, stream);
  if (rc != RTB_OK) return rc;
  ema_maps_kernel<<<g, 256, 0, stream>>>(a);
  RTB_LAUNCH_CHECK();
  rc = plain_inclusive_scan<AffineOp>(L.maps, nrows, L.tile_a, stream);
  if (rc != RTB_OK) return rc;
  ema_store_kernel<<<g, 256, 0, stream>>>(a);
  RTB_LAUNCH_CHECK();
  prof_end(PROF_SEGSCAN, stream, (uint64_t)nrows);
  return RTB_OK;
}

extern "C" int rtb_make_ifirst(const void* ikey, int kdt, int64_t nrows, int64_t n, const uint8_t* filter, int mode,
                               int32_t* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(is_ikey_dtype(kdt), "MakeiFirst: key must be int8/16/32/64");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && n >= 0 && n < (1ll << 31) - 1 && out, "MakeiFirst: bad arguments");
  RTB_ARG(mode == 0 || mode == 1, "MakeiFirst: mode must be 0 (first) or 1 (last)");
  RTB_CUDA(cudaMemsetAsync(out, mode == 0 ? 0xFF : 0x00, sizeof(int32_t) * (size_t)(n + 1), stream));
  if (nrows > 0) {
    RTB_ARG(ikey, "MakeiFirst: null key");
    ifirst_scan_kernel<<<grid_for(nrows, 256, 4), 256, 0, stream>>>(ikey, kdt, nrows, n, filter, mode, (uint32_t*)out);
    RTB_LAUNCH_CHECK();
  }
  ifirst_finish_kernel<<<grid_for(n + 1, 256, 1), 256, 0, stream>>>((uint32_t*)out, n + 1, mode);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" size_t rtb_make_inext_workspace_bytes(int64_t nrows, int64_t n) {
  (void)n;
  if (nrows < 0) return 0;
  return pack_layout(nullptr, 0, nrows, 1).bytes + 256;
}

extern "C" int rtb_make_inext(const void* ikey, int kdt, int64_t nrows, int64_t n, int mode, int32_t* out, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(is_ikey_dtype(kdt), "MakeiNext: key must be int8/16/32/64");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && n >= 0 && n < (1ll << 31) - 1, "MakeiNext: bad sizes");
  RTB_ARG(mode == 0 || mode == 1, "MakeiNext: mode must be 0 (next) or 1 (previous)");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(ikey && out, "MakeiNext: null pointer");
  RTB_ARG(((uintptr_t)ikey % vec4_align(dtype_size(kdt))) == 0, "MakeiNext: key must be aligned to 4 rows");
  PackLayout L = pack_layout(workspace, workspace_bytes, nrows, 1);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "MakeiNext: workspace too small");
  uint32_t *skeys, *ig;
  int rc = pack_sort(ikey, kdt, nrows, n + 1, L, stream, &skeys, &ig, "MakeiNext");
  if (rc != RTB_OK) return rc;
  inext_kernel<<<grid_for(nrows, 256, 2), 256, 0, stream>>>(skeys, ig, nrows, mode, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}
