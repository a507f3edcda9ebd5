// 8f.4: the row-sized steps of riptable's join-index construction (riptable/rt_merge.py:611-975, 734-1190) on device.
//
// riptable builds the two fancy indices of a join from the two sides' groupings: every left row is repeated once per
// right row that carries the same key (`_fused_arange_repeat` / `_partial_repeat`, rt_merge.py:887-975), and beside it
// go those right rows, read group by group out of the right side's packed layout (`_build_right_fancyindex_impl`,
// rt_merge.py:611-665).  Three passes here:
//   rtb_valid_mask          which rows have a usable key (`_create_column_valid_mask`, rt_merge.py:1848-1872)
//   rtb_merge_row_counts    per left row: the right key id it maps to (`_fancy_index_fetch_keymap`, rt_merge.py:668-731)
//                           and how many output rows it produces
//   rtb_merge_expand        per left row: write its run of (left row, right row) pairs at its offset
// The exclusive scan of the counts in between is the library's device-wide scan.
#include "common.cuh"
#include "values.cuh"

namespace rtb {

constexpr int32_t kInvalidIndex = (int32_t)0x80000000;  // riptable's int32 invalid (rt_enum.py:88-143)

// mode 0: out = valid; 1: out &= valid; 2: out |= valid
__global__ void __launch_bounds__(256) valid_mask_kernel(const void* __restrict__ col, int dt, int64_t n, int mode,
                                                         uint8_t* __restrict__ out) {
  const int sz = dsize(dt);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const uint8_t v = raw_invalid(load_raw1(col, sz, r), dt) ? 0 : 1;
    out[r] = mode == 0 ? v : (mode == 1 ? (uint8_t)(out[r] & v) : (uint8_t)(out[r] | v));
  }
}

__global__ void __launch_bounds__(256) merge_row_counts_kernel(const int32_t* __restrict__ left_ikey, int64_t n,
                                                               const int32_t* __restrict__ keymap, int64_t nmap,
                                                               const int32_t* __restrict__ right_ncount, int64_t ngroups,
                                                               uint32_t nonmatched, uint32_t invalid_rows,
                                                               int32_t* __restrict__ keyid_out, uint32_t* __restrict__ count_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t lk = left_ikey[i];
    int32_t keyid = kInvalidIndex;
    if (lk > 0 && lk <= nmap) keyid = __ldg(keymap + lk - 1);  // 1-based right key id, or negative
    uint32_t c = lk == 0 ? invalid_rows : nonmatched;
    if (keyid > 0 && keyid < ngroups) c = (uint32_t)__ldg(right_ncount + keyid);
    else keyid = kInvalidIndex;
    keyid_out[i] = keyid;
    count_out[i] = c;
  }
}

// 64-bit total of the counts: the scan below runs in 32 bits, and a join whose result does not fit an int32 row index has
// to be refused rather than wrapped.
__global__ void __launch_bounds__(256) count_total_kernel(const uint32_t* __restrict__ counts, int64_t n,
                                                          unsigned long long* __restrict__ total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long t = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) t += counts[i];
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0 && t) atomicAdd(total, t);
}

// One warp per 32 left rows.  Short runs are written by their own lane; a long run is taken over by the whole warp,
// one after the other, so a key with a million matches is copied with coalesced loads and stores.
__global__ void __launch_bounds__(256) merge_expand_kernel(const int32_t* __restrict__ keyid, const uint32_t* __restrict__ offset,
                                                           const uint32_t* __restrict__ count, int64_t n,
                                                           const int32_t* __restrict__ left_rows,
                                                           const int32_t* __restrict__ right_ifirstgroup,
                                                           const int32_t* __restrict__ right_igroup, int32_t* __restrict__ left_out,
                                                           int32_t* __restrict__ right_out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = ((((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5) * 32; base < n; base += nwarps * 32) {
    const int64_t i = base + lane;
    uint32_t c = 0, off = 0;
    int32_t k = kInvalidIndex, lrow = 0, g0 = 0;
    if (i < n) {
      c = count[i];
      off = offset[i];
      k = keyid[i];
      lrow = left_rows ? left_rows[i] : (int32_t)i;
      if (k > 0) g0 = __ldg(right_ifirstgroup + k);
    }
    constexpr uint32_t kShort = 8;
    if (c <= kShort) {
      for (uint32_t j = 0; j < c; j++) {
        if (left_out) left_out[off + j] = lrow;
        if (right_out) right_out[off + j] = k > 0 ? right_igroup[g0 + j] : kInvalidIndex;
      }
    }
    uint32_t longs = __ballot_sync(0xffffffffu, c > kShort);
    while (longs) {
      const int src = __ffs(longs) - 1;
      longs &= longs - 1;
      const uint32_t cc = __shfl_sync(0xffffffffu, c, src), oo = __shfl_sync(0xffffffffu, off, src);
      const int32_t kk = __shfl_sync(0xffffffffu, k, src), ll = __shfl_sync(0xffffffffu, lrow, src),
                    gg = __shfl_sync(0xffffffffu, g0, src);
      for (uint32_t j = lane; j < cc; j += 32) {
        if (left_out) left_out[oo + j] = ll;
        if (right_out) right_out[oo + j] = kk > 0 ? right_igroup[gg + j] : kInvalidIndex;
      }
    }
  }
}

}  // namespace rtb

using namespace rtb;

extern "C" int rtb_valid_mask(const void* col, int dtype, int64_t n, int mode, uint8_t* out, void* stream) {
  RTB_ARG(dtype >= RTB_BOOL && dtype <= RTB_F64, "valid mask: a numeric column");
  RTB_ARG(n >= 0 && mode >= 0 && mode <= 2, "valid mask: bad arguments");
  if (n == 0) return RTB_OK;
  RTB_ARG(col && out, "valid mask: null pointer");
  valid_mask_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(col, dtype, n, mode, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_merge_row_counts(const int32_t* left_ikey, int64_t n, const int32_t* keymap, int64_t nmap,
                                    const int32_t* right_ncountgroup, int64_t ngroups, int64_t nonmatched_count,
                                    int64_t invalid_row_count, int32_t* keyid_out, uint32_t* count_out, void* stream) {
  RTB_ARG(n >= 0 && n <= 2147483646ll && nmap >= 0 && ngroups >= 0, "merge row counts: bad sizes");
  RTB_ARG((nonmatched_count == 0 || nonmatched_count == 1) && (invalid_row_count == 0 || invalid_row_count == 1),
          "merge row counts: nonmatched_count / invalid_row_count must be 0 or 1");
  if (n == 0) return RTB_OK;
  RTB_ARG(left_ikey && keyid_out && count_out && (keymap || nmap == 0) && (right_ncountgroup || ngroups == 0),
          "merge row counts: null pointer");
  merge_row_counts_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(left_ikey, n, keymap, nmap, right_ncountgroup, ngroups,
                                                                                 (uint32_t)nonmatched_count, (uint32_t)invalid_row_count,
                                                                                 keyid_out, count_out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" size_t rtb_merge_scan_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return sizeof(uint32_t) * (scan_u32_workspace_elems(n) + 8) + 256;
}

// offsets_out[i] = sum of counts[j], j < i; *total_host = sum of all (the number of output rows).  Synchronises.
extern "C" int rtb_merge_scan_counts(const uint32_t* counts, int64_t n, uint32_t* offsets_out, int64_t* total_host,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(n >= 0 && total_host, "merge scan: bad arguments");
  *total_host = 0;
  if (n == 0) return RTB_OK;
  RTB_ARG(counts && offsets_out && workspace && workspace_bytes >= rtb_merge_scan_workspace_bytes(n) - 256 &&
              ((uintptr_t)workspace % 16) == 0,
          "merge scan: workspace too small");
  uint32_t* total_dev = (uint32_t*)workspace;
  unsigned long long* wide_dev = (unsigned long long*)(total_dev + 4);
  uint32_t* scratch = total_dev + 8;
  RTB_CUDA(cudaMemsetAsync(wide_dev, 0, sizeof(*wide_dev), stream));
  count_total_kernel<<<grid_for(n, 256, 8, 4), 256, 0, stream>>>(counts, n, wide_dev);
  RTB_LAUNCH_CHECK();
  int rc = exclusive_scan_u32(counts, offsets_out, n, SCAN_IDENTITY, scratch, total_dev, stream);
  if (rc != RTB_OK) return rc;
  unsigned long long h = 0;
  RTB_CUDA(cudaMemcpyAsync(&h, wide_dev, sizeof(h), cudaMemcpyDeviceToHost, stream));
  RTB_CUDA(cudaStreamSynchronize(stream));
  *total_host = (int64_t)h;
  RTB_ARG(h <= 2147483646ull, "merge scan: the join has %llu rows, more than an int32 row index can address", h);
  return RTB_OK;
}

extern "C" int rtb_merge_expand(const int32_t* keyid, const uint32_t* offsets, const uint32_t* counts, int64_t n,
                                const int32_t* left_rows, const int32_t* right_ifirstgroup, const int32_t* right_igroup,
                                int32_t* left_out, int32_t* right_out, void* stream) {
  RTB_ARG(n >= 0 && n <= 2147483646ll, "merge expand: bad sizes");
  if (n == 0) return RTB_OK;
  RTB_ARG(keyid && offsets && counts && (left_out || right_out), "merge expand: null pointer");
  RTB_ARG(!right_out || (right_ifirstgroup && right_igroup), "merge expand: the right side's packed layout is missing");
  merge_expand_kernel<<<grid_for(n, 256, 1, 8), 256, 0, (cudaStream_t)stream>>>(keyid, offsets, counts, n, left_rows, right_ifirstgroup,
                                                                                right_igroup, left_out, right_out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}
