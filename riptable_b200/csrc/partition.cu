// `cutoffs=` for the sort / pack entry points: rc.LexSort32(..., cutoffs=), rc.GroupFromLexSort(..., cutoffs=) and
// rc.GroupByPack32(ikey, None, n, cutoffs=) (riptable/rt_numpy.py:2199-2207, 1972).
//
// `cutoffs` (int64 cumulative row ends) splits the rows into contiguous partitions that behave as arrays of their own
// (the semantics riptable's tests pin for the hash case, riptable/tests/test_pgroupby.py:7-86).  Instead of running
// one small sort per partition, the partition id rides along as one more, most significant, key column: one stable
// sort (or pack) over all rows then holds every partition's result in its own contiguous range of positions, and
// these helpers translate between global and partition-relative numbers.  None of them moves more than one
// streaming pass over the rows.
#include "common.cuh"

namespace rtb {

// partition of a row: index of the first cutoff > row
__device__ __forceinline__ int32_t find_partition(const int64_t* __restrict__ cutoffs, int ncut, int64_t r) {
  int lo = 0, hi = ncut;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cutoffs + mid) > r) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Rows are visited in order, so a thread's rows almost always share its first row's partition: one binary search per
// thread, then a forward walk.
__global__ void __launch_bounds__(256) partition_ids_kernel(int32_t* __restrict__ pid, int64_t n,
                                                            const int64_t* __restrict__ cutoffs, int ncut) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; r < n; r += stride) {
    int32_t p = find_partition(cutoffs, ncut, r);
    int32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      while (p < ncut - 1 && __ldg(cutoffs + p) <= r + i) p++;
      o[i] = p;
    }
    if (r + 3 < n) {
      *(int4*)(pid + r) = make_int4(o[0], o[1], o[2], o[3]);
    } else {
      for (int i = 0; r + i < n; i++) pid[r + i] = o[i];
    }
  }
}

// values[i] += sign * (first row of the partition that POSITION i lies in)
__global__ void __launch_bounds__(256) partition_shift_kernel(int32_t* __restrict__ v, int64_t n, const int32_t* __restrict__ pid,
                                                              const int64_t* __restrict__ cutoffs, int sign) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t p = pid[i];
    const int32_t start = p ? (int32_t)__ldg(cutoffs + p - 1) : 0;
    v[i] += sign * start;
  }
}

// per-partition tallies over a table of rows (first rows of bins): count[pid[row]] += 1
__global__ void __launch_bounds__(256) partition_count_rows_kernel(const int32_t* __restrict__ rows, int64_t m,
                                                                   const int32_t* __restrict__ pid, int32_t* __restrict__ count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < m; b += stride) atomicAdd(&count[pid[rows[b]]], 1);
}

// iKey numbered over all partitions -> numbering that restarts in every partition (bins are partition-major, so
// partition p owns the contiguous bin range (base[p], base[p + 1]]); first rows -> partition-relative
__global__ void __launch_bounds__(256) partition_localize_kernel(int32_t* __restrict__ ikey, int64_t n, const int32_t* __restrict__ pid,
                                                                 const int32_t* __restrict__ bin_base, int32_t shift) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int32_t v = ikey[r];
    if (v > 0) ikey[r] = v - __ldg(bin_base + pid[r]) - shift;
  }
}
__global__ void __launch_bounds__(256) partition_relative_rows_kernel(int32_t* __restrict__ rows, int64_t m,
                                                                      const int32_t* __restrict__ pid,
                                                                      const int64_t* __restrict__ cutoffs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < m; b += stride) {
    const int32_t row = rows[b];
    const int32_t p = pid[row];
    rows[b] = row - (p ? (int32_t)__ldg(cutoffs + p - 1) : 0);
  }
}

// largest iKey of every partition
__global__ void __launch_bounds__(256) partition_max_ikey_kernel(const void* __restrict__ ikey, int kdt, int64_t n,
                                                                 const int32_t* __restrict__ pid, int32_t* __restrict__ umax,
                                                                 uint32_t* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t k = load_ikey(ikey, kdt, r);
    if (k < 0 || k > 0x7ffffff0ll) {
      atomicAdd(bad, 1u);
      continue;
    }
    const int32_t p = pid[r];
    if ((int32_t)k > __ldg(umax + p)) atomicMax(&umax[p], (int32_t)k);
  }
}
// composite bin over all partitions: bin_off[partition] + iKey
__global__ void __launch_bounds__(256) partition_compose_kernel(const void* __restrict__ ikey, int kdt, int64_t n,
                                                                const int32_t* __restrict__ pid, const int64_t* __restrict__ bin_off,
                                                                int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
    out[r] = (int32_t)(__ldg(bin_off + pid[r]) + load_ikey(ikey, kdt, r));
}

}  // namespace rtb

using namespace rtb;

#define RTB_PART_ARGS(who)                                                                                   \
  RTB_ARG(n >= 0 && n <= 2147483646ll && ncutoffs >= 1 && ncutoffs <= 2147483646ll, who ": bad sizes");       \
  RTB_ARG(cutoffs, who ": null cutoffs");

extern "C" int rtb_partition_ids(int64_t n, const int64_t* cutoffs, int64_t ncutoffs, int32_t* pid_out, void* stream) {
  RTB_PART_ARGS("partition ids");
  if (n == 0) return RTB_OK;
  RTB_ARG(pid_out && ((uintptr_t)pid_out % 16) == 0, "partition ids: output must be 16-byte aligned");
  partition_ids_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(pid_out, n, cutoffs, (int)ncutoffs);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_partition_shift(int32_t* values, int64_t n, const int32_t* pid, const int64_t* cutoffs, int64_t ncutoffs,
                                   int to_relative, void* stream) {
  RTB_PART_ARGS("partition shift");
  if (n == 0) return RTB_OK;
  RTB_ARG(values && pid, "partition shift: null pointer");
  partition_shift_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(values, n, pid, cutoffs, to_relative ? -1 : 1);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_partition_count_rows(const int32_t* rows, int64_t m, const int32_t* pid, int64_t ncutoffs, int32_t* count_out,
                                        void* stream) {
  RTB_ARG(m >= 0 && ncutoffs >= 1 && count_out, "partition count: bad arguments");
  RTB_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t) * (size_t)ncutoffs, (cudaStream_t)stream));
  if (m == 0) return RTB_OK;
  RTB_ARG(rows && pid, "partition count: null pointer");
  partition_count_rows_kernel<<<grid_for(m, 256, 1), 256, 0, (cudaStream_t)stream>>>(rows, m, pid, count_out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_partition_localize(int32_t* ikey, int64_t n, int32_t* ifirstkey, int64_t nbins, const int32_t* pid,
                                      const int32_t* bin_base, const int64_t* cutoffs, int64_t ncutoffs, int base_index,
                                      void* stream) {
  RTB_PART_ARGS("partition localize");
  RTB_ARG(nbins >= 0 && (base_index == 0 || base_index == 1), "partition localize: bad arguments");
  if (n == 0) return RTB_OK;
  RTB_ARG(ikey && pid && bin_base, "partition localize: null pointer");
  partition_localize_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(ikey, n, pid, bin_base, 1 - base_index);
  RTB_LAUNCH_CHECK();
  if (nbins > 0 && ifirstkey) {
    partition_relative_rows_kernel<<<grid_for(nbins, 256, 1), 256, 0, (cudaStream_t)stream>>>(ifirstkey, nbins, pid, cutoffs);
    RTB_LAUNCH_CHECK();
  }
  return RTB_OK;
}

extern "C" int rtb_partition_max_ikey(const void* ikey, int kdtype, int64_t n, const int32_t* pid, int64_t ncutoffs,
                                      int32_t* umax_out, uint32_t* bad_out, void* stream) {
  RTB_ARG(is_ikey_dtype(kdtype), "GroupByPack32: ikey must be int8/16/32/64");
  RTB_ARG(n >= 0 && ncutoffs >= 1 && umax_out && bad_out, "GroupByPack32: bad arguments");
  RTB_CUDA(cudaMemsetAsync(umax_out, 0, sizeof(int32_t) * (size_t)ncutoffs, (cudaStream_t)stream));
  RTB_CUDA(cudaMemsetAsync(bad_out, 0, sizeof(uint32_t), (cudaStream_t)stream));
  if (n == 0) return RTB_OK;
  RTB_ARG(ikey && pid, "GroupByPack32: null pointer");
  partition_max_ikey_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(ikey, kdtype, n, pid, umax_out, bad_out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}

extern "C" int rtb_partition_compose_ikey(const void* ikey, int kdtype, int64_t n, const int32_t* pid, const int64_t* bin_off,
                                          int32_t* out, void* stream) {
  RTB_ARG(is_ikey_dtype(kdtype), "GroupByPack32: ikey must be int8/16/32/64");
  RTB_ARG(n >= 0, "GroupByPack32: bad sizes");
  if (n == 0) return RTB_OK;
  RTB_ARG(ikey && pid && bin_off && out, "GroupByPack32: null pointer");
  partition_compose_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(ikey, kdtype, n, pid, bin_off, out);
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}
