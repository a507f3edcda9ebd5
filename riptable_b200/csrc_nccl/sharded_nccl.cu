// Multi-GPU C ABI (include/riptable_b200_nccl.h): the seeded, row-sharded group-and-sum step over an ncclComm_t.
// Orchestration only -- every row-sized step is an entry point of libriptable_b200.so.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/riptable_b200_nccl.h"

namespace {

constexpr int64_t kDefaultSeedRows = 1ll << 25;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#define RTBN_CUDA(call)                      \
  do {                                       \
    if ((call) != cudaSuccess) {             \
      cudaGetLastError();                    \
      return RTB_ERR_CUDA;                   \
    }                                        \
  } while (0)
#define RTBN_NCCL(call)                      \
  do {                                       \
    if ((call) != ncclSuccess) return RTB_ERR_CUDA; \
  } while (0)

// out[j] = the itemsize-byte item of `col` at row rows[j]; first64[j] = rows[j]
__global__ void __launch_bounds__(256) seed_gather_kernel(const uint8_t* __restrict__ col, int isz, const int32_t* __restrict__ rows,
                                                          int64_t n, uint8_t* __restrict__ out, int64_t* __restrict__ first64) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows[j];
    first64[j] = r;
    const uint8_t* src = col + r * isz;
    uint8_t* dst = out + j * isz;
    switch (isz) {
      case 8: *(unsigned long long*)dst = *(const unsigned long long*)src; break;
      case 4: *(uint32_t*)dst = *(const uint32_t*)src; break;
      case 2: *(uint16_t*)dst = *(const uint16_t*)src; break;
      default: *dst = *src;
    }
  }
}

struct Layout {
  uint8_t* chunk_ws;
  size_t chunk_bytes;
  int32_t* ifirst32;
  int32_t* map;
  uint8_t* seed_keys;
  int64_t* seed_first;
  int64_t* scalars;  // [0] header, [8 .. 8 + nranks) late counts
  size_t bytes;
};

Layout make_layout(void* ws, int64_t max_shard_rows, int64_t max_unique, int isz) {
  Layout L;
  size_t off = 0;
  uint8_t* base = (uint8_t*)ws;
  auto take = [&](size_t nbytes) {
    off = align_up(off, 256);
    uint8_t* p = base ? base + off : nullptr;
    off += nbytes;
    return p;
  };
  L.chunk_bytes = rtb_groupby_sum_chunk_workspace_bytes(max_shard_rows > 0 ? max_shard_rows : 1, max_unique);
  L.chunk_ws = take(L.chunk_bytes);
  L.ifirst32 = (int32_t*)take(sizeof(int32_t) * ((size_t)max_unique + 4));
  L.map = (int32_t*)take(sizeof(int32_t) * ((size_t)max_unique + 4));
  L.seed_keys = take((size_t)max_unique * isz + 16);
  L.seed_first = (int64_t*)take(sizeof(int64_t) * ((size_t)max_unique + 2));
  L.scalars = (int64_t*)take(sizeof(int64_t) * 1024);
  L.bytes = align_up(off, 256);
  return L;
}

}  // namespace

extern "C" size_t rtb_sharded_groupby_sum32_workspace_bytes(int64_t max_shard_rows, int64_t max_unique, int key_itemsize) {
  if (max_shard_rows < 0 || max_unique < 1 || key_itemsize < 1) return 0;
  return make_layout(nullptr, max_shard_rows, max_unique, key_itemsize).bytes + 256;
}

extern "C" int rtb_sharded_groupby_sum32_nccl(ncclComm_t comm, int rank, int nranks, const rtb_column* key, const void* values,
                                              int vdtype, int func, int64_t nrows, int64_t seed_rows, int32_t* ikey_out,
                                              int64_t* ifirstkey_out, void* sums_out, int64_t max_unique,
                                              rtb_sharded_result* result_host, void* workspace, size_t workspace_bytes,
                                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!comm || !key || !result_host || !sums_out || !ifirstkey_out || nranks < 1 || nranks > 1000 || rank < 0 || rank >= nranks ||
      nrows < 0 || nrows > 2147483646ll || max_unique < 1)
    return RTB_ERR_ARG;
  const int isz = key->itemsize;
  if (!(isz == 1 || isz == 2 || isz == 4 || isz == 8)) return RTB_ERR_ARG;
  if (nrows > 0 && (!key->data || !values || !ikey_out)) return RTB_ERR_ARG;
  memset(result_host, 0, sizeof(*result_host));
  Layout L = make_layout(workspace, nrows, max_unique, isz);
  if (!workspace || ((uintptr_t)workspace % 256) != 0 || L.bytes > workspace_bytes) return RTB_ERR_ARG;
  if (seed_rows <= 0) seed_rows = kDefaultSeedRows;
  const int64_t S = nrows > seed_rows ? (seed_rows & ~127ll) : nrows;  // this rank's prefix
  const int vsz = (vdtype == RTB_F64 || vdtype == RTB_I64 || vdtype == RTB_U64) ? 8 : (vdtype == RTB_F32 || vdtype == RTB_I32 || vdtype == RTB_U32) ? 4
                  : (vdtype == RTB_I16 || vdtype == RTB_U16) ? 2 : 1;

  // 1. every rank groups (and sums) its own prefix; shard 0's is the seed
  rtb_gb_state st;
  memset(&st, 0, sizeof(st));
  int64_t needed = 0;
  int rc = rtb_multikey_groupby_sum32_chunk(key, S, nullptr, values, vdtype, func, 1, ikey_out, L.ifirst32, max_unique, sums_out, 0, &st,
                                            0, &needed, L.chunk_ws, L.chunk_bytes, stream);
  bool failed = rc != RTB_OK;

  // 2. header (seed size, or -1 if shard 0 failed), then the seed itself: key items followed by first rows
  int64_t header = rank == 0 ? (failed ? -1 : st.unique) : 0;
  RTBN_CUDA(cudaMemcpyAsync(L.scalars, &header, sizeof(header), cudaMemcpyHostToDevice, stream));
  RTBN_NCCL(ncclBroadcast(L.scalars, L.scalars, 1, ncclInt64, 0, comm, stream));
  RTBN_CUDA(cudaMemcpyAsync(&header, L.scalars, sizeof(header), cudaMemcpyDeviceToHost, stream));
  RTBN_CUDA(cudaStreamSynchronize(stream));
  const int64_t u0 = header;
  if (u0 < 0) return RTB_ERR_WORKSPACE;  // every rank leaves here
  if (u0 > 0) {
    if (rank == 0) {
      int grid = (int)((u0 + 255) / 256);
      if (grid > 148 * 16) grid = 148 * 16;
      seed_gather_kernel<<<grid, 256, 0, stream>>>((const uint8_t*)key->data, isz, L.ifirst32, u0, L.seed_keys, L.seed_first);
      if (cudaGetLastError() != cudaSuccess) return RTB_ERR_CUDA;
    }
    RTBN_NCCL(ncclBroadcast(L.seed_keys, L.seed_keys, (size_t)u0 * isz, ncclUint8, 0, comm, stream));
    RTBN_NCCL(ncclBroadcast(L.seed_first, L.seed_first, (size_t)u0, ncclInt64, 0, comm, stream));
  }

  // 3. the other ranks adopt the seed's bins in place and re-index their prefix
  if (rank != 0 && !failed) {
    if (S > 0 && u0 > 0) {
      rtb_column seed{L.seed_keys, key->dtype, isz};
      int64_t nlate = 0;
      const int64_t u_loc = st.unique;  // bins of the local numbering: the map has u_loc + 1 entries
      rc = rtb_multikey_groupby32_adopt_seed(&seed, u0, L.ifirst32, max_unique, sums_out, &st, L.map, &nlate, L.chunk_ws, L.chunk_bytes,
                                             stream);
      if (rc == RTB_OK) rc = rtb_remap_ikey(ikey_out, S, L.map, u_loc + 1, ikey_out, stream);  // in place
      failed = rc != RTB_OK;
    } else if (u0 > 0) {
      // no local prefix: load the seed into a fresh table as the first chunk
      rtb_column seed{L.seed_keys, key->dtype, isz};
      memset(&st, 0, sizeof(st));
      rc = rtb_multikey_groupby_sum32_chunk(&seed, u0, nullptr, nullptr, vdtype, func, 1, L.map, L.ifirst32, max_unique, sums_out, 0, &st, 0,
                                            &needed, L.chunk_ws, L.chunk_bytes, stream);
      failed = rc != RTB_OK || st.unique != u0;
    }
  }

  // 4. the rest of the shard, hashed against the (adopted) table
  if (!failed && nrows > S) {
    rtb_column rest{(const uint8_t*)key->data + S * isz, key->dtype, isz};
    rc = rtb_multikey_groupby_sum32_chunk(&rest, nrows - S, nullptr, (const uint8_t*)values + S * vsz, vdtype, func, 1, ikey_out + S,
                                          L.ifirst32, max_unique, sums_out, S, &st, 1, &needed, L.chunk_ws, L.chunk_bytes, stream);
    failed = rc != RTB_OK;
  }

  // 5. late keys / failures, agreed on by every rank
  int64_t late = failed ? -1 : st.unique - u0;
  int64_t* lates_dev = L.scalars + 8;
  RTBN_CUDA(cudaMemcpyAsync(L.scalars + 1, &late, sizeof(late), cudaMemcpyHostToDevice, stream));
  RTBN_NCCL(ncclAllGather(L.scalars + 1, lates_dev, 1, ncclInt64, comm, stream));
  int64_t lates[1000];
  RTBN_CUDA(cudaMemcpyAsync(lates, lates_dev, sizeof(int64_t) * nranks, cudaMemcpyDeviceToHost, stream));
  RTBN_CUDA(cudaStreamSynchronize(stream));
  int64_t total_late = 0;
  bool any_failed = false;
  for (int r = 0; r < nranks; r++) {
    if (lates[r] < 0) any_failed = true;
    else total_late += lates[r];
  }
  result_host->seed_rows = seed_rows;
  result_host->late_keys = total_late;
  result_host->unique_count = u0;
  if (any_failed) return RTB_ERR_WORKSPACE;
  if (total_late > 0) return RTB_SHARDED_LATE_KEYS;

  // 6. the partial sums are indexed by global bin already: one all-reduce; first rows come from the seed
  const bool isfloat = vdtype == RTB_F32 || vdtype == RTB_F64;
  RTBN_NCCL(ncclAllReduce(sums_out, sums_out, (size_t)u0 + 1, isfloat ? ncclDouble : ncclInt64, ncclSum, comm, stream));
  if (u0 > 0) RTBN_CUDA(cudaMemcpyAsync(ifirstkey_out, L.seed_first, sizeof(int64_t) * (size_t)u0, cudaMemcpyDeviceToDevice, stream));
  RTBN_CUDA(cudaStreamSynchronize(stream));
  return RTB_OK;
}
