// 8f.2: per-group quantiles (rc.GroupByAllPack32, GB_MEDIAN 103 and GB_QUANTILE_MULT 106 —
// riptable/rt_enum.py:506-509, riptable/rt_groupbyops.py:1560-1731, 2498-2513, 1846-1861).
//
// The rows are ordered by (bin, value) with two stable LSD radix sorts — first on an order-preserving code of the
// value (invalid values code above every valid one), then on the bin — so that the run of each bin in the final
// order is the packed run of iFirstGroup / nCountGroup with its values ascending and its NaN / sentinel rows last.
// One thread per bin then reads the two order statistics around (valid - 1) * q: riptable interpolates with numpy's
// "midpoint" rule, (lower + higher) / 2 (riptable/rt_numpy.py:4153-4245).
#include "common.cuh"
#include "radix.cuh"
#include "values.cuh"

namespace rtb {

// Order-preserving sort code; invalid values get a code above every valid one.  Dtypes narrower than 8 bytes stay
// below 2^34 so that the high digits are constant and their radix passes are skipped.
__device__ __forceinline__ unsigned long long quantile_invalid_code(int vdt) {
  return dsize(vdt) < 8 ? (1ull << 33) : ~0ull;
}
__device__ __forceinline__ unsigned long long quantile_code(unsigned long long raw, int vdt) {
  switch (vdt) {
    case RTB_F64: return (raw >> 63) ? ~raw : (raw | 0x8000000000000000ull);  // +inf < ~0
    case RTB_F32: {
      uint32_t u = (uint32_t)raw;
      return (u >> 31) ? (uint32_t)~u : (u | 0x80000000u);
    }
    case RTB_I64: return (raw ^ 0x8000000000000000ull) - 1;  // the sentinel (code 0 before the shift) never gets here
    case RTB_I8: case RTB_I16: case RTB_I32: return (unsigned long long)(raw_to_i64(raw, vdt) + (1ll << 31));
    default: return raw;  // bool and the unsigned types; uint64's sentinel is its largest value
  }
}

__global__ void __launch_bounds__(256) quantile_codes_kernel(const void* __restrict__ values, int vdt, int64_t n,
                                                             unsigned long long* __restrict__ codes) {
  const int vsz = dsize(vdt);
  const unsigned long long inv = quantile_invalid_code(vdt);
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long raw = load_raw1(values, vsz, r);
    codes[r] = raw_invalid(raw, vdt) ? inv : quantile_code(raw, vdt);
  }
}

__global__ void __launch_bounds__(256) quantile_bins_kernel(const void* __restrict__ ikey, int kdt, const uint32_t* __restrict__ rows,
                                                            int64_t n, int64_t nb, uint32_t* __restrict__ bins) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = load_ikey(ikey, kdt, rows[j]);
    bins[j] = (b < 0 || b >= nb) ? 0u : (uint32_t)b;
  }
}

__global__ void __launch_bounds__(256) quantile_pick_kernel(const void* __restrict__ values, int vdt, const uint32_t* __restrict__ rows,
                                                            const int32_t* __restrict__ ifirst, const int32_t* __restrict__ ncount,
                                                            int64_t nb, double q, int nan_aware, int64_t lo, int64_t hi,
                                                            double* __restrict__ out) {
  const int vsz = dsize(vdt);
  const double nan = __longlong_as_double(0x7FF8000000000000ll);
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    double r = nan;
    const int64_t start = ifirst[b], cnt = ncount[b];
    if (b >= lo && b < hi && cnt > 0) {
      // valid rows lead the run: binary search for the first invalid one
      int64_t a = 0, z = cnt;
      while (a < z) {
        const int64_t m = (a + z) >> 1;
        if (raw_invalid(load_raw1(values, vsz, rows[start + m]), vdt)) z = m;
        else a = m + 1;
      }
      const int64_t valid = a;
      if (valid > 0 && (nan_aware || valid == cnt)) {
        const double idx = (double)(valid - 1) * q;  // numpy's virtual index, no snapping (np.quantile does none either)
        const int64_t il = (int64_t)floor(idx), ih = (int64_t)ceil(idx);
        const double vl = raw_to_f64(load_raw1(values, vsz, rows[start + il]), vdt);
        const double vh = ih == il ? vl : raw_to_f64(load_raw1(values, vsz, rows[start + ih]), vdt);
        r = (vl + vh) / 2;
      }
    }
    out[b] = r;
  }
}

// GB_MODE: the most frequent valid value of each bin; among equally frequent values the smallest (the first longest run
// of the ascending order).  One thread walks its bin's run.
__global__ void __launch_bounds__(256) mode_pick_kernel(const void* __restrict__ values, int vdt, const uint32_t* __restrict__ rows,
                                                        const int32_t* __restrict__ ifirst, const int32_t* __restrict__ ncount,
                                                        int64_t nb, int64_t lo, int64_t hi, void* __restrict__ out) {
  const int vsz = dsize(vdt);
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long best = invalid_raw(vdt);
    const int64_t start = ifirst[b], cnt = ncount[b];
    if (b >= lo && b < hi && cnt > 0) {
      int64_t best_len = 0, run_len = 0;
      unsigned long long run_raw = 0;
      for (int64_t i = 0; i < cnt; i++) {
        const unsigned long long raw = load_raw1(values, vsz, rows[start + i]);
        if (raw_invalid(raw, vdt)) break;  // invalid values close the run of the bin
        if (run_len > 0 && raw == run_raw) {
          run_len++;
        } else {
          run_raw = raw;
          run_len = 1;
        }
        if (run_len > best_len) {
          best_len = run_len;
          best = run_raw;
        }
      }
    }
    store_raw(out, vdt, b, best);
  }
}

struct QuantLayout {
  unsigned long long *codes_a, *codes_b;
  uint32_t *rows_a, *rows_b, *bins_a, *bins_b;
  void* radix_ws;
  size_t radix_bytes, bytes;
};
static QuantLayout quant_layout(void* ws, size_t wsbytes, int64_t n) {
  Arena a(ws, wsbytes);
  QuantLayout L;
  const size_t m = (size_t)(n > 0 ? n : 1);
  L.codes_a = a.take<unsigned long long>(m);
  L.codes_b = a.take<unsigned long long>(m);
  L.rows_a = a.take<uint32_t>(m);
  L.rows_b = a.take<uint32_t>(m);
  // the second sort reuses the code buffers for its bins
  L.bins_a = (uint32_t*)L.codes_a;
  L.bins_b = (uint32_t*)L.codes_b;
  L.radix_bytes = radix_workspace_bytes(n);
  L.radix_ws = a.take<uint8_t>(L.radix_bytes);
  L.bytes = a.off;
  return L;
}

static int bits_for_bins(int64_t nbins) {
  int b = 1;
  while (b < 32 && (1ll << b) < nbins) b++;
  return b;
}

}  // namespace rtb

using namespace rtb;

extern "C" size_t rtb_quantile_workspace_bytes(int64_t nrows, int64_t unique_rows) {
  (void)unique_rows;
  if (nrows < 0) return 0;
  return quant_layout(nullptr, 0, nrows).bytes + 256;
}

// Orders the rows by (bin, value); *rows_out points into the workspace.
static int sort_rows_by_bin_value(const void* values, int vdt, const void* ikey, int kdt, int64_t nrows, int64_t nb,
                                  const QuantLayout& L, cudaStream_t stream, const uint32_t** rows_out) {
  *rows_out = L.rows_a;
  if (nrows == 0) return RTB_OK;
  quantile_codes_kernel<<<grid_for(nrows, 256, 4), 256, 0, stream>>>(values, vdt, nrows, L.codes_a);
  RTB_LAUNCH_CHECK();
  int in_b = 0;
  const int vbits = dtype_size(vdt) < 8 ? 34 : 64;
  int rc = radix_sort_pairs<uint64_t>((uint64_t*)L.codes_a, L.rows_a, (uint64_t*)L.codes_b, L.rows_b, nrows, vbits, true,
                                      L.radix_ws, L.radix_bytes, stream, &in_b);
  if (rc != RTB_OK) return rc;
  uint32_t* by_value = in_b ? L.rows_b : L.rows_a;
  uint32_t* spare = in_b ? L.rows_a : L.rows_b;
  quantile_bins_kernel<<<grid_for(nrows, 256, 4), 256, 0, stream>>>(ikey, kdt, by_value, nrows, nb, L.bins_a);
  RTB_LAUNCH_CHECK();
  // the pair buffers of the second sort: (bins_a, by_value) <-> (bins_b, spare)
  rc = radix_sort_pairs<uint32_t>(L.bins_a, by_value, L.bins_b, spare, nrows, bits_for_bins(nb), false, L.radix_ws,
                                  L.radix_bytes, stream, &in_b);
  if (rc != RTB_OK) return rc;
  *rows_out = in_b ? spare : by_value;
  return RTB_OK;
}

extern "C" int rtb_groupby_quantile(const void* values, int vdt, const void* ikey, int kdt, const int32_t* ifirstgroup,
                                    const int32_t* ncountgroup, int64_t nrows, int64_t unique_rows, double q, int nan_aware,
                                    int64_t bin_low, int64_t bin_high, double* out, void* workspace, size_t workspace_bytes,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(is_ikey_dtype(kdt), "GroupByAllPack32: ikey must be int8/16/32/64");
  if (vdt < RTB_BOOL || vdt > RTB_F64) {
    set_error("GroupByAllPack32: quantile is not provided for dtype %d", vdt);
    return RTB_ERR_UNSUPPORTED;
  }
  RTB_ARG(q >= 0.0 && q <= 1.0, "GroupByAllPack32: quantile must lie in [0, 1]");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && unique_rows >= 0 && out && ifirstgroup && ncountgroup, "GroupByAllPack32: bad arguments");
  RTB_ARG(nrows == 0 || (values && ikey), "GroupByAllPack32: null pointer");
  const int64_t nb = unique_rows + 1;
  QuantLayout L = quant_layout(workspace, workspace_bytes, nrows);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "GroupByAllPack32: workspace too small");
  const uint32_t* rows = nullptr;
  int rc = sort_rows_by_bin_value(values, vdt, ikey, kdt, nrows, nb, L, stream, &rows);
  if (rc != RTB_OK) return rc;
  quantile_pick_kernel<<<grid_for(nb, 256, 1), 256, 0, stream>>>(values, vdt, rows, ifirstgroup, ncountgroup, nb, q, nan_aware,
                                                                 bin_low, bin_high, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_groupby_mode(const void* values, int vdt, const void* ikey, int kdt, const int32_t* ifirstgroup,
                                const int32_t* ncountgroup, int64_t nrows, int64_t unique_rows, int64_t bin_low,
                                int64_t bin_high, void* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(is_ikey_dtype(kdt), "GroupByAllPack32: ikey must be int8/16/32/64");
  if (vdt < RTB_BOOL || vdt > RTB_F64) {
    set_error("GroupByAllPack32: mode is not provided for dtype %d", vdt);
    return RTB_ERR_UNSUPPORTED;
  }
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll && unique_rows >= 0 && out && ifirstgroup && ncountgroup, "GroupByAllPack32: bad arguments");
  RTB_ARG(nrows == 0 || (values && ikey), "GroupByAllPack32: null pointer");
  const int64_t nb = unique_rows + 1;
  QuantLayout L = quant_layout(workspace, workspace_bytes, nrows);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "GroupByAllPack32: workspace too small");
  const uint32_t* rows = nullptr;
  int rc = sort_rows_by_bin_value(values, vdt, ikey, kdt, nrows, nb, L, stream, &rows);
  if (rc != RTB_OK) return rc;
  mode_pick_kernel<<<grid_for(nb, 256, 1), 256, 0, stream>>>(values, vdt, rows, ifirstgroup, ncountgroup, nb, bin_low, bin_high, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}
