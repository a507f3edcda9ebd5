// Sort-free GB_CUMSUM / GB_CUM[NAN]MIN / GB_CUM[NAN]MAX for groupings of up to 16 Ki bins (cumsum_blocked.cu).
#pragma once
#include "common.cuh"

namespace rtb {

struct CbPlan {
  bool ok;             // the blocked form applies to this call
  bool warp_chunks;    // few bins: a chunk per warp (no table, no turns); otherwise a chunk per CTA
  int chunks;
  int64_t chunk_rows;  // multiple of the tile size
  int64_t stride;      // accumulators per chunk (U + 1 rounded up to 16)
  size_t bytes;        // workspace: chunks x stride x 8
};

// Whether rtb_groupby_cumop(func, ...) over nrows rows and unique_rows bins takes the blocked form, and its geometry.
CbPlan cumsum_blocked_plan(int func, int64_t nrows, int64_t unique_rows, bool has_reset);

int cumsum_blocked_run(const CbPlan& p, int func, const void* values, int vdt, int odt, const void* ikey, int kdt, int64_t nrows,
                       int64_t unique_rows, const uint8_t* filter, void* out, void* workspace, cudaStream_t stream);

}  // namespace rtb
