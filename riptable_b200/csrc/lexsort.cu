// a11 / a12: rc.LexSort32 (riptable/rt_numpy.py:2658-2673, 2200-2202) and rc.GroupFromLexSort (rt_numpy.py:2207).
//
// LexSort = np.lexsort: stable, LAST key primary, NaN last.  Implemented as an LSD sort over the key columns
// from the first (least significant) to the last: each column is gathered through the current permutation
// into order-preserving unsigned codes (signed -> offset binary, floats -> monotone bits with every NaN
// mapped to the maximum and -0.0 folded onto +0.0, fixed-width strings -> big-endian 8-byte chunks, UCS4 ->
// code-point pairs) and stably radix-sorted (radix.cu).  Descending keys invert the code (NaN still last).
#include "common.cuh"
#include "keycols.cuh"
#include "radix.cuh"
#include "values.cuh"

namespace rtb {

struct SortCol {
  const uint8_t* data;
  int dtype, itemsize, chunk;  // chunk: which 8-byte (BYTES) / 2-code-point (UCS4) piece
  int nvalid;                  // bytes (BYTES, 1..8) / code points (UCS4, 1..2) the piece really holds
  int descending;
};

__device__ __forceinline__ unsigned long long sort_code(const SortCol& c, int64_t row) {
  const uint8_t* p = c.data + row * (int64_t)c.itemsize;
  unsigned long long code, width_mask;
  switch (c.dtype) {
    case RTB_BOOL: case RTB_U8: code = *p; width_mask = 0xFFull; break;
    case RTB_I8: code = (unsigned long long)(*p ^ 0x80u); width_mask = 0xFFull; break;
    case RTB_U16: code = *(const uint16_t*)p; width_mask = 0xFFFFull; break;
    case RTB_I16: code = (unsigned long long)(*(const uint16_t*)p ^ 0x8000u); width_mask = 0xFFFFull; break;
    case RTB_U32: code = *(const uint32_t*)p; width_mask = 0xFFFFFFFFull; break;
    case RTB_I32: code = (unsigned long long)(*(const uint32_t*)p ^ 0x80000000u); width_mask = 0xFFFFFFFFull; break;
    case RTB_U64: code = *(const unsigned long long*)p; width_mask = ~0ull; break;
    case RTB_I64: code = *(const unsigned long long*)p ^ 0x8000000000000000ull; width_mask = ~0ull; break;
    case RTB_F32: {
      uint32_t u = *(const uint32_t*)p;
      if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFull;  // NaN: last, ascending or not
      if (u == 0x80000000u) u = 0;                                 // -0.0 == +0.0
      code = (u >> 31) ? (uint32_t)~u : (u | 0x80000000u);
      width_mask = 0xFFFFFFFFull;
      break;
    }
    case RTB_F64: {
      unsigned long long u = *(const unsigned long long*)p;
      if ((u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return ~0ull;
      if (u == 0x8000000000000000ull) u = 0;
      code = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
      width_mask = ~0ull;
      break;
    }
    case RTB_UCS4: {
      // piece = two code points, first one most significant.  Uniform branches only (see the note on BYTES).
      const uint32_t* cp = (const uint32_t*)p + c.chunk * 2;
      code = (unsigned long long)cp[0] << 32;
      if (c.nvalid > 1) code |= cp[1];
      width_mask = ~0ull;
      break;
    }
    default: {  // RTB_BYTES: memcmp order = big-endian bytes, zero padded
      // A real loop with a grid-uniform trip count.  The fully unrolled form with a per-byte `offset < itemsize`
      // predicate was miscompiled by ptxas 12.9 for sm_100a (an under-stalled predicated move picked up a stale
      // register: byte 5 of S4 items came out as byte 3, intermittently) — measured on B200, see DESIGN.md.
      const uint8_t* q = p + c.chunk * 8;
      code = 0;
#pragma unroll 1
      for (int i = 0; i < c.nvalid; i++) code |= (unsigned long long)q[i] << (56 - 8 * i);
      width_mask = ~0ull;
    }
  }
  return c.descending ? (~code & width_mask) : code;
}

// codes[i] = sort_code(col, rows ? rows[i] : i)
template <typename K>
__global__ void __launch_bounds__(256) sort_gather_kernel(SortCol c, const uint32_t* __restrict__ rows, int64_t m,
                                                          K* __restrict__ codes) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
    codes[i] = (K)sort_code(c, rows ? (int64_t)rows[i] : i);
}

// ---- sample-sort partition: destination rank of every row given G-1 splitters ------------------------------------
constexpr int kMaxSortCols = 16;
struct LexPartArgs {
  SortCol cols[kMaxSortCols];   // local key columns, list order (first = least significant)
  SortCol split[kMaxSortCols];  // the splitter rows: same dtypes, nsplit rows each
  int ncols, nsplit;
  const long long* split_row;   // global row id of each splitter (the final tie-break)
  long long row_offset, n;
  int32_t* dest;
};

__device__ __forceinline__ int sort_pieces(const SortCol& c) {
  if (c.dtype == RTB_BYTES) return (c.itemsize + 7) / 8;
  if (c.dtype == RTB_UCS4) return (c.itemsize / 4 + 1) / 2;
  return 1;
}
__device__ __forceinline__ void set_piece(SortCol& c, int piece) {
  c.chunk = piece;
  c.nvalid = c.dtype == RTB_BYTES ? (c.itemsize - 8 * piece < 8 ? c.itemsize - 8 * piece : 8)
             : c.dtype == RTB_UCS4 ? (c.itemsize / 4 - 2 * piece < 2 ? c.itemsize / 4 - 2 * piece : 2) : 1;
}

// dest[r] = number of splitters that sort at or before row r under (key tuple, global row)
__global__ void __launch_bounds__(256) lex_partition_kernel(LexPartArgs a) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.n; r += stride) {
    int dest = 0;
    for (int j = 0; j < a.nsplit; j++) {
      int cmp = 0;  // sign of (row - splitter)
      for (int c = a.ncols - 1; c >= 0 && cmp == 0; c--) {
        SortCol x = a.cols[c], y = a.split[c];
        const int pieces = sort_pieces(x);
        for (int p = 0; p < pieces && cmp == 0; p++) {
          set_piece(x, p);
          set_piece(y, p);
          unsigned long long cx = sort_code(x, r), cy = sort_code(y, j);
          cmp = cx < cy ? -1 : (cx > cy ? 1 : 0);
        }
      }
      if (cmp == 0) cmp = (a.row_offset + r) < a.split_row[j] ? -1 : 1;
      dest += cmp > 0;
    }
    a.dest[r] = dest;
  }
}

struct LexLayout {
  unsigned long long *codes_a, *codes_b;
  uint32_t* vals_b;
  void* radix_ws;
  size_t radix_bytes, bytes;
};
static LexLayout lex_layout(void* ws, size_t wsbytes, int64_t m) {
  Arena a(ws, wsbytes);
  LexLayout L;
  size_t mm = (size_t)(m > 0 ? m : 1);
  L.codes_a = a.take<unsigned long long>(mm);
  L.codes_b = a.take<unsigned long long>(mm);
  L.vals_b = a.take<uint32_t>(mm);
  L.radix_bytes = radix_workspace_bytes(m);
  L.radix_ws = a.take<uint8_t>(L.radix_bytes);
  L.bytes = a.off;
  return L;
}

// ---- GroupFromLexSort ---------------------------------------------------------------------------------
// flag[j] = 1 where the key row at sorted position j differs (bytewise) from the one before it
__global__ void __launch_bounds__(256) lexgroup_flags_kernel(KeyCols kc, const int32_t* __restrict__ igroup, int64_t m,
                                                             uint32_t* __restrict__ flags) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride)
    flags[j] = (j == 0 || !rows_equal(kc, igroup[j], igroup[j - 1])) ? 1u : 0u;
}

// The same for up to four columns of 1 / 2 / 4 / 8-byte items (the numeric case).  The row fetches are random DRAM accesses
// and the whole cost of this call, so as few as possible are made: columns are compared from the LAST one backwards (in
// sorted order the least significant key is the one that tells neighbours apart), a lane stops at its first difference, a
// column is fetched once per position -- the predecessor's value comes from the lane below (lane 0 fetches it) -- and only
// while this position or the next still needs it, and the warp leaves the loop when every lane has its answer.
template <int NC>
__global__ void __launch_bounds__(256) lexgroup_flags_numeric_kernel(KeyCols kc, const int32_t* __restrict__ igroup, int64_t m,
                                                                     uint32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < m; base += stride) {
    const int64_t j = base + lane;
    const bool valid = j < m;
    const int64_t row = valid ? igroup[j] : 0;
    const int64_t prow = (lane == 0 && valid && j > 0) ? igroup[j - 1] : 0;
    bool open = valid && j > 0;  // still comparing: no difference found so far
#pragma unroll
    for (int c = NC - 1; c >= 0; c--) {
      bool next_open = __shfl_down_sync(0xffffffffu, open, 1);
      if (lane == 31) next_open = false;  // the next warp's lane 0 fetches its predecessor itself
      const unsigned long long v = (valid && (open || next_open)) ? load_raw1(kc.ptr[c], kc.itemsize[c], row) : 0ull;
      unsigned long long pv = __shfl_up_sync(0xffffffffu, v, 1);
      if (lane == 0 && open) pv = load_raw1(kc.ptr[c], kc.itemsize[c], prow);
      if (open && v != pv) open = false;
      if (!__any_sync(0xffffffffu, open)) break;
    }
    // `open` still set: every column equal -> same bin as the predecessor
    if (valid) flags[j] = (j == 0 || !open) ? 1u : 0u;
  }
}

// binid[j] (exclusive scan of flags, + flag = 1-based bin) -> iKey, iFirstKey, bin start positions
__global__ void __launch_bounds__(256) lexgroup_assign_kernel(const int32_t* __restrict__ igroup, int64_t m,
                                                              const uint32_t* __restrict__ flags,
                                                              const uint32_t* __restrict__ excl, int base_index,
                                                              int32_t* __restrict__ ikey, int32_t* __restrict__ ifirstkey,
                                                              uint32_t* __restrict__ starts) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
    uint32_t f = flags[j];
    uint32_t bin = excl[j] + f;  // 1-based
    int32_t row = igroup[j];
    ikey[row] = (int32_t)bin - (1 - base_index);
    if (f) {
      ifirstkey[bin - 1] = row;
      starts[bin - 1] = (uint32_t)j;
    }
  }
}

__global__ void __launch_bounds__(256) lexgroup_counts_kernel(const uint32_t* __restrict__ starts, int64_t U, int64_t m,
                                                              int32_t* __restrict__ ncount) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= U; b += (int64_t)gridDim.x * blockDim.x) {
    if (b == 0) ncount[0] = 0;
    else ncount[b] = (int32_t)((b < U ? starts[b] : (uint32_t)m) - starts[b - 1]);
  }
}

struct LgLayout {
  uint32_t *flags, *excl, *starts, *scan_scratch, *total;
  size_t bytes;
};
static LgLayout lg_layout(void* ws, size_t wsbytes, int64_t m) {
  Arena a(ws, wsbytes);
  LgLayout L;
  size_t mm = (size_t)(m > 0 ? m : 1);
  L.flags = a.take<uint32_t>(mm);
  L.excl = a.take<uint32_t>(mm);
  L.starts = a.take<uint32_t>(mm + 1);
  L.scan_scratch = a.take<uint32_t>(scan_u32_workspace_elems(m > 0 ? m : 1));
  L.total = a.take<uint32_t>(4);
  L.bytes = a.off;
  return L;
}

}  // namespace rtb

using namespace rtb;

extern "C" size_t rtb_lexsort_workspace_bytes(int64_t nrows, int64_t nindex, int nkeys, const int32_t* itemsizes_host) {
  (void)nrows; (void)nkeys; (void)itemsizes_host;
  if (nindex < 0) return 0;
  return lex_layout(nullptr, 0, nindex).bytes + 256;
}

// The sort itself, shared by the 32- and 64-bit entry points: row ids travel as uint32 through the radix engine, so any
// row count below 2^32 - 1 sorts the same way; only the result's dtype differs.  `perm` (m uint32) receives the result
// and doubles as a working buffer; `index_u32` (or nullptr) lists the rows to sort.
static int lexsort_core(const rtb_column* keys, int nkeys, int64_t m, const uint32_t* index_u32, const int8_t* ascending_host,
                        uint32_t* perm, const LexLayout& L, cudaStream_t stream, const char* who) {
  for (int c = 0; c < nkeys; c++) {
    RTB_ARG(keys[c].data && keys[c].itemsize >= 1, "%s: bad key column %d", who, c);
    RTB_ARG(keys[c].dtype >= RTB_BOOL && keys[c].dtype <= RTB_UCS4, "%s: bad dtype for key column %d", who, c);
    RTB_ARG(keys[c].dtype > RTB_F64 || keys[c].itemsize == dtype_size(keys[c].dtype), "%s: itemsize/dtype mismatch", who);
    RTB_ARG(keys[c].dtype != RTB_UCS4 || keys[c].itemsize % 4 == 0, "%s: UCS4 itemsize must be a multiple of 4", who);
  }
  uint32_t* cur = perm;  // current permutation (valid once `started`)
  uint32_t* other = L.vals_b;
  bool started = false;
  if (index_u32) {
    if (index_u32 != cur) RTB_CUDA(cudaMemcpyAsync(cur, index_u32, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToDevice, stream));
    started = true;
  }
  for (int c = 0; c < nkeys; c++) {
    SortCol sc;
    sc.data = (const uint8_t*)keys[c].data;
    sc.dtype = keys[c].dtype;
    sc.itemsize = keys[c].itemsize;
    sc.descending = ascending_host && !ascending_host[c];
    int pieces = 1, bits = 8 * keys[c].itemsize;
    if (sc.dtype == RTB_BYTES) { pieces = (sc.itemsize + 7) / 8; bits = 64; }
    if (sc.dtype == RTB_UCS4) { pieces = (sc.itemsize / 4 + 1) / 2; bits = 64; }
    const bool wide = bits > 32;
    for (int piece = pieces - 1; piece >= 0; piece--) {
      sc.chunk = piece;
      sc.nvalid = sc.dtype == RTB_BYTES ? (sc.itemsize - 8 * piece < 8 ? sc.itemsize - 8 * piece : 8)
                  : sc.dtype == RTB_UCS4 ? (sc.itemsize / 4 - 2 * piece < 2 ? sc.itemsize / 4 - 2 * piece : 2) : 1;
      int in_b = 0, rc;
      prof_begin(PROF_GATHER, stream);
      if (wide) {
        sort_gather_kernel<unsigned long long><<<grid_for(m, 256, 2), 256, 0, stream>>>(sc, started ? cur : nullptr, m, L.codes_a);
        RTB_LAUNCH_CHECK();
        prof_end(PROF_GATHER, stream, (uint64_t)m);
        rc = radix_sort_pairs<uint64_t>((uint64_t*)L.codes_a, cur, (uint64_t*)L.codes_b, other, m, bits, !started, L.radix_ws,
                                        L.radix_bytes, stream, &in_b);
      } else {
        sort_gather_kernel<uint32_t><<<grid_for(m, 256, 2), 256, 0, stream>>>(sc, started ? cur : nullptr, m, (uint32_t*)L.codes_a);
        RTB_LAUNCH_CHECK();
        prof_end(PROF_GATHER, stream, (uint64_t)m);
        rc = radix_sort_pairs<uint32_t>((uint32_t*)L.codes_a, cur, (uint32_t*)L.codes_b, other, m, bits, !started, L.radix_ws,
                                        L.radix_bytes, stream, &in_b);
      }
      if (rc != RTB_OK) return rc;
      started = true;
      if (in_b) { uint32_t* t = cur; cur = other; other = t; }
    }
  }
  if (cur != perm) RTB_CUDA(cudaMemcpyAsync(perm, cur, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToDevice, stream));
  return RTB_OK;
}

extern "C" int rtb_lexsort32(const rtb_column* keys, int nkeys, int64_t nrows, const int32_t* index, int64_t nindex,
                             const int8_t* ascending_host, int32_t* perm_out, void* workspace, size_t workspace_bytes,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(keys && nkeys >= 1, "LexSort32: first argument is not a list of arrays");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll, "LexSort32: row count out of range for an int32 result");
  const int64_t m = index ? nindex : nrows;
  RTB_ARG(m >= 0 && m <= 2147483646ll, "LexSort32: bad index length");
  if (m == 0) return RTB_OK;
  RTB_ARG(perm_out, "LexSort32: null output");
  LexLayout L = lex_layout(workspace, workspace_bytes, m);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "LexSort32: workspace too small");
  return lexsort_core(keys, nkeys, m, (const uint32_t*)index, ascending_host, (uint32_t*)perm_out, L, stream, "LexSort32");
}

// ---- rc.LexSort64 (riptable/rt_numpy.py:2199-2200, 2669-2670: what riptable calls above 2e9 rows): int64 result, int64
// `index`.  Row ids still travel as uint32 inside the radix engine, so this entry sorts up to 2^32 - 2 rows.
__global__ void __launch_bounds__(256) widen_u32_to_i64_kernel(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int64_t)in[i];
}
__global__ void __launch_bounds__(256) narrow_i64_to_u32_kernel(const int64_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (uint32_t)in[i];
}

extern "C" size_t rtb_lexsort64_workspace_bytes(int64_t nrows, int64_t nindex, int nkeys, const int32_t* itemsizes_host) {
  (void)nrows; (void)nkeys; (void)itemsizes_host;
  if (nindex < 0) return 0;
  return align_up(lex_layout(nullptr, 0, nindex).bytes, 256) + sizeof(uint32_t) * (size_t)(nindex > 0 ? nindex : 1) + 512;
}

extern "C" int rtb_lexsort64(const rtb_column* keys, int nkeys, int64_t nrows, const int64_t* index, int64_t nindex,
                             const int8_t* ascending_host, int64_t* perm_out, void* workspace, size_t workspace_bytes,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(keys && nkeys >= 1, "LexSort64: first argument is not a list of arrays");
  RTB_ARG(nrows >= 0 && nrows <= 4294967294ll, "LexSort64: at most 2^32 - 2 rows");
  const int64_t m = index ? nindex : nrows;
  RTB_ARG(m >= 0 && m <= 4294967294ll, "LexSort64: bad index length");
  if (m == 0) return RTB_OK;
  RTB_ARG(perm_out, "LexSort64: null output");
  const size_t lex_bytes = align_up(lex_layout(nullptr, 0, m).bytes, 256);
  RTB_ARG(workspace && ((uintptr_t)workspace % 256) == 0 && lex_bytes + sizeof(uint32_t) * (size_t)m <= workspace_bytes,
          "LexSort64: workspace too small or not 256-byte aligned");
  LexLayout L = lex_layout(workspace, lex_bytes, m);
  uint32_t* perm = (uint32_t*)((uint8_t*)workspace + lex_bytes);
  if (index) {
    narrow_i64_to_u32_kernel<<<grid_for(m, 256, 4), 256, 0, stream>>>(index, m, perm);
    RTB_LAUNCH_CHECK();
  }
  int rc = lexsort_core(keys, nkeys, m, index ? perm : nullptr, ascending_host, perm, L, stream, "LexSort64");
  if (rc != RTB_OK) return rc;
  widen_u32_to_i64_kernel<<<grid_for(m, 256, 4), 256, 0, stream>>>(perm, m, perm_out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

// debugging aid (not part of the public header): the 64-bit sort codes of one column
extern "C" int rtb_debug_sort_codes(const rtb_column* col, int64_t n, int chunk, int descending, unsigned long long* out, void* stream) {
  SortCol sc;
  sc.data = (const uint8_t*)col->data; sc.dtype = col->dtype; sc.itemsize = col->itemsize; sc.chunk = chunk; sc.descending = descending;
  sc.nvalid = sc.dtype == RTB_BYTES ? (sc.itemsize - 8 * chunk < 8 ? sc.itemsize - 8 * chunk : 8)
              : sc.dtype == RTB_UCS4 ? (sc.itemsize / 4 - 2 * chunk < 2 ? sc.itemsize / 4 - 2 * chunk : 2) : 1;
  sort_gather_kernel<unsigned long long><<<grid_for(n, 256, 2), 256, 0, (cudaStream_t)stream>>>(sc, nullptr, n, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_lex_partition(const rtb_column* keys, int nkeys, int64_t nrows, const int8_t* ascending_host,
                                 const rtb_column* splitters, int nsplit, const int64_t* splitter_rows, int64_t row_offset,
                                 int32_t* dest_out, void* stream) {
  RTB_ARG(keys && nkeys >= 1 && nkeys <= kMaxSortCols, "lex partition: 1..%d key columns", kMaxSortCols);
  RTB_ARG(nsplit >= 0 && nrows >= 0, "lex partition: bad sizes");
  if (nrows == 0) return RTB_OK;
  RTB_ARG(dest_out && (nsplit == 0 || (splitters && splitter_rows)), "lex partition: null pointer");
  LexPartArgs a;
  memset(&a, 0, sizeof(a));
  for (int c = 0; c < nkeys; c++) {
    RTB_ARG(keys[c].dtype >= RTB_BOOL && keys[c].dtype <= RTB_UCS4 && keys[c].data, "lex partition: bad key column %d", c);
    RTB_ARG(nsplit == 0 || (splitters[c].dtype == keys[c].dtype && splitters[c].itemsize == keys[c].itemsize),
            "lex partition: splitter column %d does not match the key column", c);
    a.cols[c].data = (const uint8_t*)keys[c].data;
    a.cols[c].dtype = keys[c].dtype;
    a.cols[c].itemsize = keys[c].itemsize;
    a.cols[c].descending = ascending_host && !ascending_host[c];
    a.split[c] = a.cols[c];
    a.split[c].data = nsplit ? (const uint8_t*)splitters[c].data : nullptr;
  }
  a.ncols = nkeys;
  a.nsplit = nsplit;
  a.split_row = (const long long*)splitter_rows;
  a.row_offset = row_offset;
  a.n = nrows;
  a.dest = dest_out;
  lex_partition_kernel<<<grid_for(nrows, 256, 1, 16), 256, 0, (cudaStream_t)stream>>>(a);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" size_t rtb_group_from_lexsort_workspace_bytes(int64_t nrows, int64_t nindex) {
  (void)nrows;
  if (nindex < 0) return 0;
  return lg_layout(nullptr, 0, nindex).bytes + 256;
}

extern "C" int rtb_group_from_lexsort(const int32_t* igroup, int64_t nindex, const rtb_column* keys, int nkeys, int64_t nrows,
                                      int base_index, int32_t* ikey_out, int32_t* ifirstkey_out, int32_t* ncountgroup_out,
                                      int64_t* unique_count_host, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(keys && nkeys >= 1 && nkeys <= kMaxKeyCols - 1, "GroupFromLexSort: bad key list");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && nindex >= 0 && nindex <= nrows, "GroupFromLexSort: bad sizes");
  RTB_ARG(unique_count_host && ncountgroup_out, "GroupFromLexSort: null output");
  RTB_ARG(base_index == 0 || base_index == 1, "GroupFromLexSort: base_index must be 0 or 1");
  *unique_count_host = 0;
  if (nrows > 0) {
    RTB_ARG(ikey_out, "GroupFromLexSort: null iKey output");
    RTB_CUDA(cudaMemsetAsync(ikey_out, 0, sizeof(int32_t) * (size_t)nrows, stream));
  }
  RTB_CUDA(cudaMemsetAsync(ncountgroup_out, 0, sizeof(int32_t), stream));
  if (nindex == 0) return RTB_OK;
  RTB_ARG(igroup && ifirstkey_out, "GroupFromLexSort: null pointer");
  LgLayout L = lg_layout(workspace, workspace_bytes, nindex);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "GroupFromLexSort: workspace too small");
  KeyCols kc;
  RTB_ARG(make_keycols(keys, nkeys, &kc), "GroupFromLexSort: too many key columns");
  bool numeric = kc.ncols <= 4;
  for (int c = 0; c < kc.ncols && numeric; c++)
    numeric = (kc.itemsize[c] == 1 || kc.itemsize[c] == 2 || kc.itemsize[c] == 4 || kc.itemsize[c] == 8) &&
              ((uintptr_t)kc.ptr[c] % kc.itemsize[c]) == 0;
  const int fg = grid_for(nindex, 256, 2);
  if (numeric) {
    switch (kc.ncols) {
      case 1: lexgroup_flags_numeric_kernel<1><<<fg, 256, 0, stream>>>(kc, igroup, nindex, L.flags); break;
      case 2: lexgroup_flags_numeric_kernel<2><<<fg, 256, 0, stream>>>(kc, igroup, nindex, L.flags); break;
      case 3: lexgroup_flags_numeric_kernel<3><<<fg, 256, 0, stream>>>(kc, igroup, nindex, L.flags); break;
      default: lexgroup_flags_numeric_kernel<4><<<fg, 256, 0, stream>>>(kc, igroup, nindex, L.flags);
    }
  } else {
    lexgroup_flags_kernel<<<fg, 256, 0, stream>>>(kc, igroup, nindex, L.flags);
  }
  RTB_LAUNCH_CHECK();
  int rc = exclusive_scan_u32(L.flags, L.excl, nindex, SCAN_IDENTITY, L.scan_scratch, L.total, stream);
  if (rc != RTB_OK) return rc;
  lexgroup_assign_kernel<<<grid_for(nindex, 256, 2), 256, 0, stream>>>(igroup, nindex, L.flags, L.excl, base_index, ikey_out,
                                                                       ifirstkey_out, L.starts);
  RTB_LAUNCH_CHECK();
  uint32_t hU = 0;
  RTB_CUDA(cudaMemcpyAsync(&hU, L.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  RTB_CUDA(cudaStreamSynchronize(stream));
  lexgroup_counts_kernel<<<grid_for((int64_t)hU + 1, 256, 1), 256, 0, stream>>>(L.starts, hU, nindex, ncountgroup_out);
  RTB_LAUNCH_CHECK();
  *unique_count_host = hU;
  return RTB_OK;
}
