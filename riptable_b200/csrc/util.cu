// Error state, launch counter, device-wide scan, and the small streaming entry points
// (CombineFilter, ReverseShuffle, remap).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace rtb {

static thread_local char g_err[512] = "";
uint64_t g_launches = 0;
int g_debug_sync = getenv("RTB_DEBUG_SYNC") ? atoi(getenv("RTB_DEBUG_SYNC")) : 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- profiling slots ---------------------------------------------------------------------------
int g_prof_on = 0;
struct ProfPending { int id; cudaEvent_t a, b; uint64_t units; };
static ProfPending g_pending[4096];
static int g_npending = 0;
static cudaEvent_t g_open[PROF_COUNT];
static double g_prof_ms[PROF_COUNT];
static uint64_t g_prof_launches[PROF_COUNT], g_prof_units[PROF_COUNT];

static void prof_drain() {
  for (int i = 0; i < g_npending; i++) {
    ProfPending& p = g_pending[i];
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      g_prof_ms[p.id] += ms;
      g_prof_launches[p.id] += 1;
      g_prof_units[p.id] += p.units;
    }
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  g_npending = 0;
}
void prof_begin(int id, cudaStream_t s) {
  if (!g_prof_on) return;
  cudaEventCreate(&g_open[id]);
  cudaEventRecord(g_open[id], s);
}
void prof_end(int id, cudaStream_t s, uint64_t units) {
  if (!g_prof_on) return;
  if (g_npending == 4096) prof_drain();
  ProfPending& p = g_pending[g_npending++];
  p.id = id; p.a = g_open[id]; p.units = units;
  cudaEventCreate(&p.b);
  cudaEventRecord(p.b, s);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan: tile = 256 threads x 16 items
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t xform(uint32_t v, int xf) { return xf == SCAN_POPC ? (uint32_t)__popc(v) : v; }

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  // 256 threads; returns exclusive prefix of v across the block
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  __shared__ uint32_t block_total;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    uint32_t si = s;
#pragma unroll
    for (int d = 1; d < kScanThreads / 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d) si += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = si - s;
    if (lane == kScanThreads / 32 - 1) block_total = si;
  }
  __syncthreads();
  uint32_t r = inc - v + warp_sums[w];
  *total = block_total;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t* __restrict__ in, int64_t n, int xf,
                                                                uint32_t* __restrict__ tile_sums) {
  int64_t base = (int64_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    int64_t idx = base + i * kScanThreads + threadIdx.x;
    if (idx < n) s += xform(in[idx], xf);
  }
  uint32_t total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_of_sums(uint32_t* sums, int64_t m, uint32_t* total_dev) {
  // single block: exclusive scan of m tile sums in place
  uint32_t carry = 0;
  for (int64_t base = 0; base < m; base += kScanThreads) {
    int64_t idx = base + threadIdx.x;
    uint32_t v = idx < m ? sums[idx] : 0;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, &total);
    if (idx < m) sums[idx] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0 && total_dev) *total_dev = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t* in, uint32_t* out, int64_t n, int xf,
                                                            const uint32_t* __restrict__ tile_offsets) {
  // thread owns kScanItems consecutive elements so the scan order is the array order
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    int64_t idx = base + i;
    v[i] = idx < n ? xform(in[idx], xf) : 0;
    s += v[i];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
}

size_t scan_u32_workspace_elems(int64_t n) { return (size_t)((n + kScanTile - 1) / kScanTile) + 1; }

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, ScanXform xf, uint32_t* scratch,
                       uint32_t* total_dev, cudaStream_t stream) {
  if (n <= 0) {
    if (total_dev) RTB_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), stream));
    return RTB_OK;
  }
  int64_t tiles = (n + kScanTile - 1) / kScanTile;
  scan_tile_sums<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, n, (int)xf, scratch);
  RTB_LAUNCH_CHECK();
  scan_of_sums<<<1, kScanThreads, 0, stream>>>(scratch, tiles, total_dev);
  RTB_LAUNCH_CHECK();
  scan_apply<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, out, n, (int)xf, scratch);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

// ---------------------------------------------------------------------------------------------
// a4 CombineFilter: out = filter ? key : 0  (streaming, 4 rows per thread)
__global__ void __launch_bounds__(256) combine_filter_kernel(const void* __restrict__ ikey, int kdt,
                                                              const uint8_t* __restrict__ filter, int64_t n,
                                                              void* __restrict__ out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; r < n; r += stride) {
    if (r + 3 < n) {
      int64_t k[4];
      load_ikey4(ikey, kdt, r, k);
      uint32_t f = ld_stream_u32(filter + r);
#pragma unroll
      for (int i = 0; i < 4; i++) store_ikey(out, kdt, r + i, ((f >> (8 * i)) & 0xff) ? k[i] : 0);
    } else {
      for (int64_t q = r; q < n; q++) store_ikey(out, kdt, q, filter[q] ? load_ikey(ikey, kdt, q) : 0);
    }
  }
}

// Accum2 cell index: out = filter ? key2 * (unique_count1 + 1) + key1 : 0  (riptable/rt_numpy.py:1549-1597)
__global__ void __launch_bounds__(256) combine_accum2_kernel(const void* __restrict__ key1, int kdt1,
                                                              const void* __restrict__ key2, int kdt2,
                                                              const uint8_t* __restrict__ filter, int64_t n, int64_t stride1,
                                                              void* __restrict__ out, int odt) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    int64_t v = 0;
    if (!filter || filter[r]) v = load_ikey(key2, kdt2, r) * stride1 + load_ikey(key1, kdt1, r);
    store_ikey(out, odt, r, v);
  }
}

__global__ void __launch_bounds__(256) reverse_shuffle_kernel(const int32_t* __restrict__ perm, int64_t n,
                                                               int32_t* __restrict__ out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int32_t p = perm[i];
    if (p >= 0 && p < n) out[p] = (int32_t)i;
  }
}

__global__ void __launch_bounds__(256) remap_kernel(const int32_t* __restrict__ ikey, int64_t n,
                                                     const int32_t* __restrict__ map, int64_t nmap,
                                                     int32_t* __restrict__ out) {
  // 4 rows per thread through 128-bit loads/stores; the map (4 B x bins) is an L2-resident gather
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      uint4 k = ld_stream_u4(ikey + i);
      uint32_t kk[4] = {k.x, k.y, k.z, k.w}, o[4];
#pragma unroll
      for (int j = 0; j < 4; j++) o[j] = kk[j] < (uint32_t)nmap ? (uint32_t)__ldg(map + kk[j]) : 0u;
      st_stream_u4(out + i, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
      for (int64_t q = i; q < n; q++) {
        int32_t k = ikey[q];
        out[q] = (k >= 0 && k < nmap) ? map[k] : 0;
      }
    }
  }
}

// one partition of rc.ReIndexGroups: base-1 local bins -> global bins through that partition's slice of uikey
__global__ void __launch_bounds__(256) reindex_kernel(const int32_t* __restrict__ ikey, int64_t n,
                                                       const int32_t* __restrict__ uikey, int64_t nu,
                                                       int32_t* __restrict__ out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int32_t k = ikey[i];
    out[i] = (k > 0 && k <= nu) ? __ldg(uikey + (k - 1)) : 0;
  }
}

// out[r] = key[r] > 0 (rc.CombineAccum1Filter: rows of bin 0 drop out of the renumbering)
__global__ void __launch_bounds__(256) positive_mask_kernel(const void* __restrict__ key, int kdt, int64_t n, uint8_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) out[r] = load_ikey(key, kdt, r) > 0 ? 1 : 0;
}

// Multi-key / string ismember: the rows of `a` were grouped together with the rows of `b` ([b ; a], rows of b first).  A row
// of a is a member iff its bin's first row lies in b, and that first row is the index of the key's first occurrence in b.
__global__ void __launch_bounds__(256) ismember_finish_kernel(const int32_t* __restrict__ ikey_a, int64_t na,
                                                               const int32_t* __restrict__ ifirstkey, int64_t nuniq, int64_t nb,
                                                               int base_index, int idt, long long absent,
                                                               uint8_t* __restrict__ found, void* __restrict__ index_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < na; r += stride) {
    const int32_t bin = ikey_a[r];
    long long pos = -1;
    if (bin > 0 && bin <= nuniq) {
      const int32_t first = __ldg(ifirstkey + bin - 1);
      if (first < nb) pos = first;
    }
    found[r] = pos >= 0;
    const long long v = pos >= 0 ? pos + base_index : absent;
    switch (idt) {
      case RTB_I8: ((int8_t*)index_out)[r] = (int8_t)v; break;
      case RTB_I16: ((int16_t*)index_out)[r] = (int16_t)v; break;
      case RTB_I32: ((int32_t*)index_out)[r] = (int32_t)v; break;
      default: ((long long*)index_out)[r] = v;
    }
  }
}

// first row holding each code in [0, nslots) (0xFFFFFFFF = -1 where a code never occurs; other codes are ignored)
__global__ void __launch_bounds__(256) first_row_per_code_kernel(const void* __restrict__ codes, int dt, int64_t n, int64_t nslots,
                                                                  uint32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t c = load_ikey(codes, dt, r);
    if (c >= 0 && c < nslots && (uint32_t)r < __ldg(out + c)) atomicMin(&out[c], (uint32_t)r);
  }
}

// rc.IsMemberCategoricalFixup (riptable/rt_numpy.py:1302-1305): code of a's Categorical -> a's unique -> b's unique
// (on_unique) -> b's code -> first row of b holding that code.  Two gathers from U-sized tables per row.
__global__ void __launch_bounds__(256) ismember_cat_fixup_kernel(const void* __restrict__ a, int adt, int64_t na,
                                                                  const int32_t* __restrict__ on_unique, int64_t nua,
                                                                  const int32_t* __restrict__ b_first, int64_t num_unique_b,
                                                                  int a_base, int b_base, uint8_t* __restrict__ found,
                                                                  int32_t* __restrict__ index_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride) {
    const int64_t ua = load_ikey(a, adt, i) - a_base;  // filtered (0 in base 1) and sentinel codes fall out of range
    int32_t pos = (int32_t)0x80000000;
    if (ua >= 0 && ua < nua) {
      const int32_t ub = __ldg(on_unique + ua);
      if (ub >= 0 && ub < num_unique_b) pos = __ldg(b_first + ub + b_base);
    }
    found[i] = pos >= 0;
    index_out[i] = pos >= 0 ? pos : (int32_t)0x80000000;
  }
}

}  // namespace rtb

using namespace rtb;

extern "C" {

int rtb_positive_mask(const void* ikey, int kdtype, int64_t n, uint8_t* out, void* stream) {
  RTB_ARG(is_ikey_dtype(kdtype) && n >= 0, "positive mask: bad arguments");
  if (n == 0) return RTB_OK;
  RTB_ARG(ikey && out, "positive mask: null pointer");
  positive_mask_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(ikey, kdtype, n, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_ismember_finish(const int32_t* ikey_a, int64_t na, const int32_t* ifirstkey, int64_t unique_count, int64_t nb, int base_index,
                        int index_dtype, uint8_t* found_out, void* index_out, void* stream) {
  RTB_ARG(index_dtype == RTB_I8 || index_dtype == RTB_I16 || index_dtype == RTB_I32 || index_dtype == RTB_I64, "ismember: bad index dtype");
  RTB_ARG(na >= 0 && unique_count >= 0 && nb >= 0 && (base_index == 0 || base_index == 1), "ismember: bad arguments");
  if (na == 0) return RTB_OK;
  RTB_ARG(ikey_a && found_out && index_out && (ifirstkey || unique_count == 0), "ismember: null pointer");
  long long absent = 0;
  if (!base_index) absent = index_dtype == RTB_I8 ? -128ll : index_dtype == RTB_I16 ? -32768ll : index_dtype == RTB_I32 ? -2147483648ll : (long long)0x8000000000000000ull;
  ismember_finish_kernel<<<grid_for(na, 256, 4), 256, 0, (cudaStream_t)stream>>>(ikey_a, na, ifirstkey, unique_count, nb, base_index, index_dtype,
                                                                                absent, found_out, index_out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_first_row_per_code(const void* codes, int dtype, int64_t n, int64_t nslots, int32_t* out, void* stream) {
  RTB_ARG(is_ikey_dtype(dtype), "first row per code: the codes must be int8/16/32/64");
  RTB_ARG(n >= 0 && n <= 2147483646ll && nslots >= 0 && (out || nslots == 0), "first row per code: bad arguments");
  if (nslots == 0) return RTB_OK;
  RTB_CUDA(cudaMemsetAsync(out, 0xFF, sizeof(int32_t) * (size_t)nslots, (cudaStream_t)stream));
  if (n == 0) return RTB_OK;
  RTB_ARG(codes, "first row per code: null pointer");
  first_row_per_code_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(codes, dtype, n, nslots, (uint32_t*)out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_ismember_categorical_fixup(const void* a_codes, int adtype, int64_t na, const int32_t* on_unique, int64_t n_unique_a,
                                   const int32_t* b_first, int64_t num_unique_b, int a_base, int b_base, uint8_t* found_out,
                                   int32_t* index_out, void* stream) {
  RTB_ARG(is_ikey_dtype(adtype), "IsMemberCategoricalFixup: the codes must be int8/16/32/64");
  RTB_ARG(na >= 0 && n_unique_a >= 0 && num_unique_b >= 0, "IsMemberCategoricalFixup: bad sizes");
  RTB_ARG((a_base == 0 || a_base == 1) && (b_base == 0 || b_base == 1), "IsMemberCategoricalFixup: base index must be 0 or 1");
  if (na == 0) return RTB_OK;
  RTB_ARG(a_codes && found_out && index_out && (on_unique || n_unique_a == 0) && (b_first || num_unique_b == 0),
          "IsMemberCategoricalFixup: null pointer");
  ismember_cat_fixup_kernel<<<grid_for(na, 256, 4), 256, 0, (cudaStream_t)stream>>>(a_codes, adtype, na, on_unique, n_unique_a, b_first,
                                                                                   num_unique_b, a_base, b_base, found_out, index_out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_version(void) { return 100; }
const char* rtb_last_error(void) { return g_err; }
uint64_t rtb_launch_count(void) { return g_launches; }

void rtb_profile_enable(int on) { g_prof_on = on; }
void rtb_profile_reset(void) {
  prof_drain();
  for (int i = 0; i < PROF_COUNT; i++) { g_prof_ms[i] = 0; g_prof_launches[i] = 0; g_prof_units[i] = 0; }
}
int rtb_profile_get(int id, double* ms, uint64_t* launches, uint64_t* units) {
  if (id < 0 || id >= PROF_COUNT) return RTB_ERR_ARG;
  prof_drain();
  if (ms) *ms = g_prof_ms[id];
  if (launches) *launches = g_prof_launches[id];
  if (units) *units = g_prof_units[id];
  return RTB_OK;
}

int rtb_combine_filter(const void* ikey, int kdtype, const uint8_t* filter, int64_t nrows, void* out, void* stream) {
  RTB_ARG(is_ikey_dtype(kdtype), "CombineFilter: key must be int8/16/32/64");
  RTB_ARG(nrows >= 0, "CombineFilter: negative length");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(ikey && filter && out, "CombineFilter: null pointer");
  // vector path needs 4-row alignment of the base pointers
  int ksz = dtype_size(kdtype);
  RTB_ARG(((uintptr_t)ikey % vec4_align(ksz)) == 0 && ((uintptr_t)filter % 4) == 0, "CombineFilter: unaligned input");
  combine_filter_kernel<<<grid_for(nrows, 256, 4), 256, 0, (cudaStream_t)stream>>>(ikey, kdtype, filter, nrows, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_combine_accum2(const void* key1, int kdt1, const void* key2, int kdt2, const uint8_t* filter, int64_t nrows,
                       int64_t unique_count1, void* out, int odt, void* stream) {
  RTB_ARG(is_ikey_dtype(kdt1) && is_ikey_dtype(kdt2) && is_ikey_dtype(odt), "CombineAccum2Filter: keys must be int8/16/32/64");
  RTB_ARG(nrows >= 0 && unique_count1 >= 0, "CombineAccum2Filter: bad sizes");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(key1 && key2 && out, "CombineAccum2Filter: null pointer");
  combine_accum2_kernel<<<grid_for(nrows, 256, 2), 256, 0, (cudaStream_t)stream>>>(key1, kdt1, key2, kdt2, filter, nrows,
                                                                                  unique_count1 + 1, out, odt);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_reverse_shuffle(const int32_t* perm, int64_t n, int32_t* out, void* stream) {
  RTB_ARG(n >= 0, "ReverseShuffle: negative length");
  if (n == 0) return RTB_OK;
  reverse_shuffle_kernel<<<grid_for(n, 256, 1), 256, 0, (cudaStream_t)stream>>>(perm, n, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_remap_ikey(const int32_t* ikey, int64_t nrows, const int32_t* map, int64_t nmap, int32_t* out, void* stream) {
  if (nrows <= 0) return RTB_OK;
  RTB_ARG(((uintptr_t)ikey % 16) == 0 && ((uintptr_t)out % 16) == 0, "remap: buffers must be 16-byte aligned");
  remap_kernel<<<grid_for(nrows, 256, 4, 32), 256, 0, (cudaStream_t)stream>>>(ikey, nrows, map, nmap, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

int rtb_reindex_groups(const int32_t* ikey, int64_t nrows, const int32_t* uikey, int64_t nuikey,
                       const int64_t* u_cutoffs_host, const int64_t* i_cutoffs_host, int64_t ncutoffs, int32_t* out,
                       void* stream) {
  RTB_ARG(nrows >= 0 && nuikey >= 0 && ncutoffs >= 0, "ReIndexGroups: bad sizes");
  RTB_ARG(ncutoffs == 0 || (u_cutoffs_host && i_cutoffs_host), "ReIndexGroups: null cutoffs");
  int64_t ustart = 0, istart = 0;
  for (int64_t p = 0; p < ncutoffs; p++) {
    int64_t ustop = u_cutoffs_host[p], istop = i_cutoffs_host[p];
    RTB_ARG(ustop >= ustart && ustop <= nuikey && istop >= istart && istop <= nrows,
            "ReIndexGroups: cutoffs must be non-decreasing and within the arrays");
    if (istop > istart) {
      RTB_ARG(ikey && out && (uikey || ustop == ustart), "ReIndexGroups: null pointer");
      reindex_kernel<<<grid_for(istop - istart, 256, 4, 32), 256, 0, (cudaStream_t)stream>>>(
          ikey + istart, istop - istart, uikey + ustart, ustop - ustart, out + istart);
      RTB_LAUNCH_CHECK();
    }
    ustart = ustop;
    istart = istop;
  }
  if (istart < nrows) RTB_CUDA(cudaMemsetAsync(out + istart, 0, sizeof(int32_t) * (size_t)(nrows - istart), (cudaStream_t)stream));
  return RTB_OK;
}

}  // extern "C"
