// Cross-file entry into the reduction kernels (reduce.cu) for the fused group-and-sum path (hash_group.cu).
#pragma once
#include "common.cuh"

namespace rtb {

// Adds the values of n rows to per-bin accumulators: the accumulate pass of GB_SUM / GB_NANSUM without zeroing or
// finalising.  `values` / `ikey` point at the first row of the range and are 16-byte aligned; rows whose bin lies in
// [bin_lo, nb) take part; float columns accumulate as f64 bits, integer columns as 64-bit wrap-around integers.
int accum_sum_rows(const void* values, int vdt, const void* ikey, int kdt, int64_t n, int64_t nb, int64_t bin_lo,
                   int nanskip, unsigned long long* acc, cudaStream_t stream);

}  // namespace rtb
