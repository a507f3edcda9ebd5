/* riptable_b200 — multi-GPU C ABI: the row-sharded group-and-sum step over an NCCL communicator (last row;
 * SURVEY 8e "Hash grouping" + "Reductions").
 *
 * One process (or thread) per GPU calls the same function with its own shard.  This is riptable's partition scheme —
 * `cutoffs` partitions grouped independently, then merged by re-grouping the stacked uniques (`hstack_groupings`,
 * riptable/rt_grouping.py:277-413) — restructured so that the merge costs U-sized work only: the first shard publishes the
 * keys of a row prefix with their bins (the seed), every other shard adopts those bins into its own hash table, and all
 * shards hash their rows straight to GLOBAL bins while summing the value column in the same pass
 * (rtb_multikey_groupby_sum32_chunk, rtb_multikey_groupby32_adopt_seed of riptable_b200.h).  Collectives: one 8-byte and
 * one seed-sized ncclBroadcast, one ncclAllGather of a flag per rank, one ncclAllReduce of the U + 1 partial sums.
 *
 * Lives in its own shared library (libriptable_b200_nccl.so, linked against libriptable_b200.so and libnccl.so.2) so that
 * the single-GPU library carries no NCCL dependency.  riptable_b200/dist.py is the same algorithm over torch.distributed,
 * with the fall-back paths this entry leaves to its caller.
 */
#ifndef RIPTABLE_B200_NCCL_H
#define RIPTABLE_B200_NCCL_H

#include <nccl.h>

#include "riptable_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Positive status of rtb_sharded_groupby_sum32_nccl: the step completed on every rank but some keys lie outside the seed
 * (first seen after shard 0's prefix, or absent from shard 0).  ikey_out then holds provisional bins above the seed's for
 * those keys and sums_out is NOT reduced; the caller merges the late keys (riptable_b200/dist.py shows how) or repeats the
 * call with a longer seed prefix.  Every rank returns the same status. */
#define RTB_SHARDED_LATE_KEYS 1

typedef struct {
  int64_t unique_count; /* global number of bins (== seed size when the status is RTB_OK) */
  int64_t late_keys;    /* keys outside the seed, summed over the ranks */
  int64_t seed_rows;    /* rows of shard 0 that built the seed */
} rtb_sharded_result;

/* Device workspace for shards of at most max_shard_rows rows and max_unique bins (key items of key_itemsize bytes). */
size_t rtb_sharded_groupby_sum32_workspace_bytes(int64_t max_shard_rows, int64_t max_unique, int key_itemsize);

/* Row-sharded MultiKeyGroupBy32 + GroupByAll32(GB_SUM | GB_NANSUM) for ONE key column of 1, 2, 4 or 8 bytes.
 *   key / values        this rank's shard (device), nrows rows; shards are contiguous row ranges in rank order
 *   seed_rows           rows of shard 0 grouped before the other shards adopt its bins (0: the library's default, 32 Mi)
 *   ikey_out            int32[nrows]            GLOBAL base-1 bin of every local row
 *   ifirstkey_out       int64[max_unique]       global row of each bin's first appearance, replicated (first U valid)
 *   sums_out            [max_unique + 1] raw accumulators, replicated after the all-reduce: float64 for float columns,
 *                       int64 (wrap-around) for integer columns; slot 0 is not summed (bin_low = 1)
 *   result_host         host struct, identical on every rank
 * Returns RTB_OK, RTB_SHARDED_LATE_KEYS, or a negative rtb_status (the same on every rank: failures are exchanged before
 * the collective that follows them, so no rank is left waiting).  Synchronises `stream`. */
int rtb_sharded_groupby_sum32_nccl(ncclComm_t comm, int rank, int nranks, const rtb_column* key, const void* values,
                                   int vdtype, int func, int64_t nrows, int64_t seed_rows, int32_t* ikey_out,
                                   int64_t* ifirstkey_out, void* sums_out, int64_t max_unique,
                                   rtb_sharded_result* result_host, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIPTABLE_B200_NCCL_H */
