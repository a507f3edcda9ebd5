/* riptable_b200 — C ABI of the B200-native groupby / reduce / lexsort hot path.
 *
 * Every entry point replaces one riptide_cpp (`rc.*`) function that riptable calls on this path.
 * riptide_cpp is a CPython extension whose source is not in the reference tree; the interface each
 * function replaces is therefore cited as the riptable call site that defines its arguments
 * (paths relative to /root/reference).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - No allocation happens inside the library: scratch comes from `workspace` (device memory the
 *     caller allocates, e.g. with torch.empty) sized by the matching *_workspace_bytes query.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*).  Functions that return a
 *     host scalar (unique counts) synchronise that stream before returning.
 *   - Return value: RTB_OK or a negative rtb_status; rtb_last_error() gives the message.  The
 *     library never aborts the process.
 *   - dtype codes are rtb_dtype; fixed-width byte strings (numpy S/U/void) are RTB_BYTES with an
 *     explicit itemsize.
 */
#ifndef RIPTABLE_B200_H
#define RIPTABLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RTB_BOOL = 0, RTB_I8, RTB_U8, RTB_I16, RTB_U16, RTB_I32, RTB_U32, RTB_I64, RTB_U64, RTB_F32, RTB_F64,
  RTB_BYTES, /* fixed-width bytes (numpy S / void): memcmp order */
  RTB_UCS4   /* fixed-width UTF-32 (numpy U): code-point order; grouped bytewise like RTB_BYTES */
} rtb_dtype;

typedef enum {
  RTB_OK = 0,
  RTB_ERR_ARG = -1,         /* bad argument (shim raises ValueError/TypeError) */
  RTB_ERR_WORKSPACE = -2,   /* workspace / max_unique too small; *needed_unique_host says how much to retry with */
  RTB_ERR_CUDA = -3,        /* a CUDA call failed */
  RTB_ERR_UNSUPPORTED = -4  /* dtype/op combination not provided (shim returns None for that column) */
} rtb_status;

/* Op numbers: riptable/rt_enum.py:486-532 (GB_FUNCTIONS). */
typedef enum {
  RTB_GB_SUM = 0, RTB_GB_MEAN = 1, RTB_GB_MIN = 2, RTB_GB_MAX = 3, RTB_GB_VAR = 4, RTB_GB_STD = 5,
  RTB_GB_NANSUM = 50, RTB_GB_NANMEAN = 51, RTB_GB_NANMIN = 52, RTB_GB_NANMAX = 53, RTB_GB_NANVAR = 54,
  RTB_GB_NANSTD = 55,
  RTB_GB_FIRST = 100, RTB_GB_NTH = 101, RTB_GB_LAST = 102, RTB_GB_MEDIAN = 103, RTB_GB_MODE = 104, RTB_GB_QUANTILE_MULT = 106,
  RTB_GB_ROLLING_SUM = 200, RTB_GB_ROLLING_NANSUM = 201, RTB_GB_ROLLING_DIFF = 202, RTB_GB_ROLLING_SHIFT = 203,
  RTB_GB_ROLLING_COUNT = 204, RTB_GB_ROLLING_MEAN = 205, RTB_GB_ROLLING_NANMEAN = 206, RTB_GB_ROLLING_QUANTILE = 207,
  RTB_GB_CUMSUM = 300, RTB_GB_EMADECAY = 301, RTB_GB_CUMPROD = 302, RTB_GB_EMANORMAL = 304, RTB_GB_EMAWEIGHTED = 305, RTB_GB_CUMNANMAX = 306, RTB_GB_CUMNANMIN = 307, RTB_GB_CUMMAX = 308,
  RTB_GB_CUMMIN = 309
} rtb_gb_func;

/* One key / value column: `itemsize` bytes per row, rows contiguous. */
typedef struct {
  const void* data;
  int32_t dtype;    /* rtb_dtype */
  int32_t itemsize; /* bytes per row */
} rtb_column;

int rtb_version(void);
const char* rtb_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t rtb_launch_count(void);

/* Optional per-kernel timing with CUDA events on the launching stream (off by default).
 * ids: 0 hash-group round kernel, 1 reduce accumulate kernel, 2 radix histogram, 3 radix scatter,
 *      4 segmented scan, 5 key gather.  `units` = rows the timed launches processed. */
void rtb_profile_enable(int on);
void rtb_profile_reset(void);
int rtb_profile_get(int id, double* ms, uint64_t* launches, uint64_t* units);

/* ---- a1: rc.MultiKeyGroupBy32(list, hint_size, filter, hash_mode, cutoffs=)  — riptable/rt_numpy.py:2073-2075
 * Hash-groups the (multi)key rows; bins numbered by first appearance among rows passing `filter`.
 *   ikey_out      int32[nrows]  base-1 bin per row, 0 = filtered out
 *   ifirstkey_out int32[max_unique] first row of each bin (strictly increasing); first U entries valid.
 *                 With cutoffs: partition-relative rows (tests/test_pgroupby.py:15-25).
 *   cutoffs_host  int64[ncutoffs] cumulative partition ends (host memory) or NULL
 *   ucutoffs_host int64[ncutoffs] out: cumulative unique counts per partition (host), or NULL
 *   max_unique    capacity the caller provisioned (sizes the table and ifirstkey_out)
 * On RTB_ERR_WORKSPACE, *needed_unique_host holds a max_unique to retry with. */
size_t rtb_groupby_workspace_bytes(int64_t nrows, int64_t max_unique, int nkeys, int has_cutoffs);
/* Same, given the summed itemsize of the key columns: adds room for the packed key rows that multi-key / string
 * groupings of up to 64 bytes per row use.  With the smaller workspace above such keys take the generic kernel. */
size_t rtb_groupby_workspace_bytes_rows(int64_t nrows, int64_t max_unique, int64_t row_bytes, int has_cutoffs);
int rtb_multikey_groupby32(const rtb_column* keys, int nkeys, int64_t nrows, int64_t hint_size,
                           const uint8_t* filter, const int64_t* cutoffs_host, int64_t ncutoffs,
                           int32_t* ikey_out, int32_t* ifirstkey_out, int64_t max_unique,
                           int64_t* unique_count_host, int64_t* ucutoffs_host, int64_t* needed_unique_host,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Resumable form of a1 for ONE key column of 1, 2, 4 or 8 bytes: consecutive calls group consecutive row chunks as
 * if they were one array (bins keep counting in first-appearance order; the hash table stays in `workspace`, whose
 * size is rtb_groupby_workspace_bytes(largest chunk, max_unique, 1, 0)).  This is the `cutoffs`-free half of
 * riptable's partition merge (riptable/rt_grouping.py:277-413) done in place, and what lets a host-mode caller
 * overlap PCIe transfers with hashing.
 *   ikey_out      int32[nrows]       bins of this chunk's rows
 *   ifirstkey_out int32[max_unique]  the SAME array on every call; first rows of new bins are row_offset + local row
 *   state         host struct carried between calls (zeroed by the library when resume == 0); state->unique = U so far
 * A chunk that needs more than max_unique bins fails with RTB_ERR_WORKSPACE and the stream cannot be continued. */
typedef struct {
  int64_t unique;     /* bins found so far */
  int64_t prev_new;   /* new bins / rows of the last round: the discovery rate the next chunk starts from */
  int64_t prev_rows;
  int64_t max_unique;
  uint32_t cap;       /* table capacity in slots */
  int32_t rounds;
} rtb_gb_state;
int rtb_multikey_groupby32_chunk(const rtb_column* key, int64_t nrows, const uint8_t* filter, int32_t* ikey_out,
                                 int32_t* ifirstkey_out, int64_t max_unique, int64_t row_offset, rtb_gb_state* state,
                                 int resume, int64_t* needed_unique_host, void* workspace, size_t workspace_bytes,
                                 void* stream);

/* ---- a1 + a5 fused: rc.MultiKeyGroupBy32(keys, hint, filter) followed by rc.GroupByAll32([values], iKey, U, [GB_SUM |
 * GB_NANSUM], [bin_low], [U + 1]) — the pair of calls `ds.gb(k).sum()` makes (riptable/rt_groupbyops.py:1161-1225 ->
 * riptable/rt_grouping.py:3395-3436 -> riptable/rt_numpy.py:2073, 1871) — as ONE pass: the steady-state probe kernel adds
 * each row's value to its bin as it writes iKey, so the reduction's second sweep (iKey re-read + a second kernel) is gone.
 * Outputs are exactly those of the two calls: ikey_out / ifirstkey_out / *unique_count_host as rtb_multikey_groupby32,
 * sum_out[0 .. U] as rtb_groupby_reduce (dtype rtb_reduce_out_dtype(func, vdtype); slot 0 = filtered rows, summed only
 * when bin_low == 0).  sum_out must hold max_unique + 1 items.  Any key layout is accepted; rounds whose kernel has no
 * fused form (discovery rounds, shared-memory tables, multi-key rows) sum their rows right after they are grouped.
 * riptable binding: the lazy `GroupBy.grouping` (riptable/rt_groupby.py) lets `GroupByOps.sum` ask for the grouping and
 * the sums together when the grouping has not been computed yet. */
size_t rtb_groupby_sum_workspace_bytes(int64_t nrows, int64_t max_unique, int64_t row_bytes);
int rtb_multikey_groupby_sum32(const rtb_column* keys, int nkeys, int64_t nrows, const uint8_t* filter,
                               const void* values, int vdtype, int func, int64_t bin_low, int32_t* ikey_out,
                               int32_t* ifirstkey_out, int64_t max_unique, void* sum_out, int64_t* unique_count_host,
                               int64_t* needed_unique_host, void* workspace, size_t workspace_bytes, void* stream);

/* Resumable form of the fused call (one key column of 1, 2, 4 or 8 bytes), as rtb_multikey_groupby32_chunk is to
 * rtb_multikey_groupby32: consecutive calls group-and-sum consecutive row chunks as if they were one array.  `acc` holds
 * max_unique + 1 raw accumulators that persist between calls -- float64 for float columns, int64 (wrap-around) for
 * integer columns, indexed by bin, slot 0 = filtered rows (summed only when bin_low == 0); the library zeroes an
 * accumulator when its bin comes into existence.  `values` may be NULL: the chunk then only creates bins (this is how a
 * row shard pre-loads the keys another shard found first, riptable_b200/dist.py).  The workspace
 * (rtb_groupby_sum_chunk_workspace_bytes(largest chunk, max_unique)) must be the same buffer on every call. */
size_t rtb_groupby_sum_chunk_workspace_bytes(int64_t max_chunk_rows, int64_t max_unique);
int rtb_multikey_groupby_sum32_chunk(const rtb_column* key, int64_t nrows, const uint8_t* filter, const void* values,
                                     int vdtype, int func, int64_t bin_low, int32_t* ikey_out, int32_t* ifirstkey_out,
                                     int64_t max_unique, void* acc, int64_t row_offset, rtb_gb_state* state, int resume,
                                     int64_t* needed_unique_host, void* workspace, size_t workspace_bytes, void* stream);

/* Row shards (riptable_b200/dist.py): re-label a resumable grouping with a SEED, in place.  The grouping has so far numbered
 * its keys locally (state->unique bins); `seed_keys` holds nseed distinct keys whose bins are 1..nseed in seed order (the bins
 * the first shard gave them).  Afterwards seed keys have their seed bin (those the table lacked are inserted), the local keys
 * the seed lacks are numbered nseed + 1 .. in their local first-appearance order (*nlate_host of them), iFirstKey and the
 * accumulators `acc` (NULL, or the raw accumulators of rtb_multikey_groupby_sum32_chunk) are permuted accordingly (seed bins
 * without a local row get iFirstKey -1), and map_out[old bin] = new bin (state->unique + 1 entries, map_out[0] = 0) lets the
 * caller re-index the iKey it already holds (rtb_remap_ikey).  RTB_ERR_UNSUPPORTED: the table has no room for nseed more keys
 * (the caller then builds a fresh table from the seed instead).  Same workspace as the chunk calls. */
int rtb_multikey_groupby32_adopt_seed(const rtb_column* seed_keys, int64_t nseed, int32_t* ifirstkey, int64_t max_unique,
                                      void* acc, rtb_gb_state* state, int32_t* map_out, int64_t* nlate_host, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* ---- 8f.1: rc.IsMember32(a, b, h, hint) / rc.IsMemberCategorical — riptable/rt_numpy.py:1388-1391, one numeric key
 * column of the same dtype on both sides.  found_out uint8[na]; index_out (index_dtype: an int rtb_dtype chosen by
 * len(b), riptable/rt_enum.py:666-679) = position of the first occurrence of a[i] in b, + base_index; absent rows get
 * the index dtype's invalid (base_index 0) or 0 (base_index 1).  NaN never matches.  The table is built by grouping b
 * (a1's kernels); the rows of a are looked up without inserting.  RTB_ERR_UNSUPPORTED for an unaligned key column. */
size_t rtb_ismember_workspace_bytes(int64_t nb);
int rtb_ismember32(const void* a, int64_t na, const void* b, int64_t nb, int dtype, int base_index, int index_dtype,
                   uint8_t* found_out, void* index_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- 8f.1: rc.IsMemberCategoricalFixup(a_codes, b_codes, on_unique, num_unique_b, a_base, b_base) —
 * riptable/rt_numpy.py:1302-1305: membership of one Categorical's rows in another, from the match of their UNIQUES.
 * on_unique int32[n_unique_a] = position of a's unique in b's uniques (negative = absent); b_first int32[num_unique_b +
 * b_base] = first row of b holding each CODE of b (rtb_first_row_per_code; negative = that code never occurs).
 * found_out uint8[na]; index_out int32[na] = first row of b with the same category, int32 invalid where there is none
 * (filtered / invalid codes of a, categories missing from b).  Pinned by riptable/tests/test_ismember.py:162-195, 491-505. */
int rtb_ismember_categorical_fixup(const void* a_codes, int adtype, int64_t na, const int32_t* on_unique, int64_t n_unique_a,
                                   const int32_t* b_first, int64_t num_unique_b, int a_base, int b_base, uint8_t* found_out,
                                   int32_t* index_out, void* stream);
/* out[c] = first row whose code is c, for c in [0, nslots); -1 where a code never occurs (other codes are ignored). */
int rtb_first_row_per_code(const void* codes, int dtype, int64_t n, int64_t nslots, int32_t* out, void* stream);

/* Row-sized helpers of the shims that used to be eager tensor expressions:
 *   rtb_positive_mask    out[r] = key[r] > 0 — rc.CombineAccum1Filter (riptable/rt_numpy.py:1512-1545) drops bin 0 before it
 *                        renumbers the bins with rtb_multikey_groupby32
 *   rtb_ismember_finish  multi-key / string rc.MultiKeyIsMember32 (riptable/rt_numpy.py:1325): after grouping [b ; a] as one
 *                        array, a row of a is a member iff its bin's first row lies in b (< nb), and that row is the answer;
 *                        absent rows get the index dtype's invalid (base_index 0) or 0 (base_index 1) */
int rtb_positive_mask(const void* ikey, int kdtype, int64_t n, uint8_t* out, void* stream);
int rtb_ismember_finish(const int32_t* ikey_a, int64_t na, const int32_t* ifirstkey, int64_t unique_count, int64_t nb,
                        int base_index, int index_dtype, uint8_t* found_out, void* index_out, void* stream);

/* ---- a4: rc.CombineFilter(key, filter) — riptable/rt_numpy.py:1508.  out = filter ? key : 0, key dtype kept. */
int rtb_combine_filter(const void* ikey, int kdtype, const uint8_t* filter, int64_t nrows, void* out, void* stream);

/* ---- rc.CombineAccum2Filter(key1, key2, unique_count1, unique_count2, filter) — riptable/rt_numpy.py:1595
 * Accum2 cell index: out = filter ? key2 * (unique_count1 + 1) + key1 : 0, stored as odt (an iKey dtype). */
int rtb_combine_accum2(const void* key1, int kdt1, const void* key2, int kdt2, const uint8_t* filter, int64_t nrows,
                       int64_t unique_count1, void* out, int odt, void* stream);

/* ---- a5: rc.GroupByAll32(values, ikey, unique_rows, funcList, binLowList, binHighList, func_param)
 *          — riptable/rt_numpy.py:1871, dispatched from riptable/rt_grouping.py:3434-3436.
 * One value column per call.  out has unique_rows+1 entries of rtb_reduce_out_dtype(func, vdtype);
 * slot 0 is the filtered bin; bins outside [bin_low, bin_high) get the op's empty value. */
int rtb_reduce_out_dtype(int func, int vdtype); /* rtb_dtype, or <0 if the column is not reducible */
size_t rtb_groupby_reduce_workspace_bytes(int64_t unique_rows, int func, int vdtype);
int rtb_groupby_reduce(const void* values, int vdtype, const void* ikey, int kdtype, int64_t nrows,
                       int64_t unique_rows, int func, int64_t bin_low, int64_t bin_high, void* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- a6: rc.BinCount(ikey, n) — riptable/rt_grouping.py:1680,3532; riptable/rt_numpy.py:1956.  out int32[nbins]. */
int rtb_bincount(const void* ikey, int kdtype, int64_t nrows, int64_t nbins, int32_t* out, void* stream);

/* ---- a7: rc.BinCount(ikey, n, pack=True) / rc.GroupFromBinCount(ikey, nCountGroup) — riptable/rt_numpy.py:1957-1965
 * Stable bucket sort of rows by bin.  igroup_out int32[nrows] (rows ascending inside a bin, bin 0 first),
 * ifirstgroup_out int32[nbins] = exclusive cumsum of ncountgroup, ncountgroup_out int32[nbins]. */
size_t rtb_pack_workspace_bytes(int64_t nrows, int64_t nbins);
int rtb_groupby_pack(const void* ikey, int kdtype, int64_t nrows, int64_t nbins, int32_t* igroup_out,
                     int32_t* ifirstgroup_out, int32_t* ncountgroup_out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---- a8: rc.GroupByAllPack32(values, ikey, iGroup, iFirstGroup, nCountGroup, unique_rows, funcList, binLow,
 *          binHigh, func_param) — riptable/rt_numpy.py:1891, riptable/rt_grouping.py:3443-3454.
 * GB_FIRST / GB_NTH / GB_LAST; out has unique_rows+1 items of the value's own dtype/itemsize. */
int rtb_groupby_pick(const void* values, int vdtype, int32_t itemsize, const int32_t* igroup,
                     const int32_t* ifirstgroup, const int32_t* ncountgroup, int64_t unique_rows, int func,
                     int64_t nth, int64_t bin_low, int64_t bin_high, void* out, void* stream);

/* ---- a9: rc.EmaAll32(values, ikey, unique_rows, GB_CUMSUM, (decay, time, filter, reset_filter))
 *          — riptable/rt_grouping.py:3430.  Per-group running sum in row order.
 * Two forms: without a reset filter, up to 4 Mi bins and from 4 Mi rows, a blocked scan over the rows in their own order
 * (per-chunk bin totals, a scan over the chunks, then every chunk's rows in order against a shared-memory table of
 * running values; workspace = chunks x bins x 8 bytes, <= 8 GB); otherwise through the packed layout (sort by bin, gather,
 * segmented scan, scatter).  rtb_cumop_workspace_bytes says what one call needs, rtb_cumsum_workspace_bytes is enough
 * for either form.
 * out: nrows items of rtb_cumsum_out_dtype(vdtype); bin-0 rows get the invalid of that dtype. */
int rtb_cumsum_out_dtype(int vdtype);
size_t rtb_cumsum_workspace_bytes(int64_t nrows, int64_t unique_rows);
size_t rtb_cumop_workspace_bytes(int func, int64_t nrows, int64_t unique_rows, int has_reset_filter);
int rtb_groupby_cumsum(const void* values, int vdtype, const void* ikey, int kdtype, int64_t nrows,
                       int64_t unique_rows, const uint8_t* filter, const uint8_t* reset_filter, void* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- 8f.2: the other cumulative functions of rc.EmaAll32 — riptable/rt_groupbyops.py:3182-3258, function numbers
 * riptable/rt_enum.py (GB_CUMPROD 302, GB_CUMNANMAX 306, GB_CUMNANMIN 307, GB_CUMMAX 308, GB_CUMMIN 309).
 * func = RTB_GB_CUM*.  cumprod: out dtype as cumsum.  cummin / cummax: out dtype = vdtype
 * (riptable/tests/test_groupbyops.py:524-525); the NAN variants skip NaN / integer sentinels (np.fmin.accumulate),
 * the plain ones turn the rest of the group invalid after the first one (np.minimum.accumulate); a row before the
 * group's first valid value gets the invalid.  Workspace: rtb_cumsum_workspace_bytes (or rtb_cumop_workspace_bytes for
 * exactly one call).  cummin / cummax take the sort-free blocked form under the same conditions as cumsum. */
int rtb_cumop_out_dtype(int func, int vdtype);
int rtb_groupby_cumop(int func, const void* values, int vdtype, const void* ikey, int kdtype, int64_t nrows,
                      int64_t unique_rows, const uint8_t* filter, const uint8_t* reset_filter, void* out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- 8f.2: the rolling functions of rc.GroupByAllPack32 (function numbers 200-206, riptable/rt_enum.py:513-519;
 * riptable/rt_groupbyops.py:2941-3153, 3500-3670).  Row-shaped output (nrows items) in original row order; bin 0 is
 * computed as a group of its own (riptable/rt_grouping.py:3324-3326) except for ROLLING_COUNT (cumcount), where it gets
 * the invalid.  `window`: the window length for SUM / NANSUM / MEAN / NANMEAN (partial windows at the start of a group
 * use what is there, riptable/tests/test_fastarray.py:1297-1305), the period for SHIFT / DIFF (may be negative), the
 * direction for COUNT (>= 0 ascending, < 0 descending).  Output dtype: rtb_rolling_out_dtype. */
int rtb_rolling_out_dtype(int func, int vdtype);
int rtb_groupby_rolling(int func, const void* values, int vdtype, int32_t itemsize, const void* ikey, int kdtype,
                        const int32_t* igroup, const int32_t* ifirstgroup, const int32_t* ncountgroup, int64_t nrows,
                        int64_t unique_rows, int64_t window, void* out, void* stream);

/* ---- 8f.2: per-group quantiles — rc.GroupByAllPack32 with GB_MEDIAN (103) / GB_QUANTILE_MULT (106),
 * riptable/rt_groupbyops.py:1560-1731 (func_param = int(q * 1e9) + is_nan * (1e9 + 1)), 2498-2513 (median = q 0.5).
 * out: unique_rows+1 float64; numpy "midpoint" rule on the group's valid values (riptable/rt_numpy.py:4153-4245);
 * nan_aware = 0: a group holding a NaN / integer sentinel gives NaN.  Bins outside [bin_low, bin_high) and empty
 * groups give NaN.  iFirstGroup / nCountGroup are the packed layout's (rtb_groupby_pack). */
size_t rtb_quantile_workspace_bytes(int64_t nrows, int64_t unique_rows);
int rtb_groupby_quantile(const void* values, int vdtype, const void* ikey, int kdtype, const int32_t* ifirstgroup,
                         const int32_t* ncountgroup, int64_t nrows, int64_t unique_rows, double q, int nan_aware,
                         int64_t bin_low, int64_t bin_high, double* out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- 8f.2: rc.EmaAll32 time-decayed averages — GB_EMADECAY (301), GB_EMANORMAL (304), GB_EMAWEIGHTED (305),
 * func_param = (decay_rate, time, filter, reset_filter); formulas in riptable/rt_groupbyops.py:3306-3498:
 *   decay:    out = x + last * exp(-rate * (t - t_last))
 *   normal:   w = exp(-rate * (t - t_last)); out = x * (1 - w) + last * w      (a group starts at its first value)
 *   weighted: out = x * (1 - rate) + last * rate                               (no time array)
 * Rows the op-filter excludes show the group's current value (0 before the first included row) and do not advance
 * t_last; reset_filter restarts the group at that row; bin-0 rows get the invalid.  out: float32 for float32 values,
 * float64 otherwise (rtb_ema_out_dtype).  time: any numeric dtype (tdtype). */
int rtb_ema_out_dtype(int func, int vdtype);
size_t rtb_ema_workspace_bytes(int64_t nrows, int64_t unique_rows);
int rtb_groupby_ema(int func, const void* values, int vdtype, const void* ikey, int kdtype, int64_t nrows,
                    int64_t unique_rows, const void* time, int tdtype, double decay_rate, const uint8_t* filter,
                    const uint8_t* reset_filter, void* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- 8f.2: rc.GroupByAllPack32 GB_MODE (104, "auto handles nan", riptable/rt_enum.py:507; riptable/rt_groupbyops.py:1297-1363).
 * out: unique_rows+1 items of the value dtype: the most frequent valid value of each bin, the smallest one among
 * equally frequent values; a bin with no valid value, or outside [bin_low, bin_high), gets the invalid.
 * Workspace: rtb_quantile_workspace_bytes. */
int rtb_groupby_mode(const void* values, int vdtype, const void* ikey, int kdtype, const int32_t* ifirstgroup,
                     const int32_t* ncountgroup, int64_t nrows, int64_t unique_rows, int64_t bin_low, int64_t bin_high,
                     void* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- 8f.2: rolling nan-quantile — rc.GroupByAllPack32 function 207, func_param = int(q * 1e9) + window * (1e9 + 1)
 * (riptable/rt_utils.py:1172-1184; riptable/rt_groupbyops.py:2997-3080).  out: nrows float64 in original row order: the
 * midpoint quantile of the valid values among the last `window` rows of the row's group (fewer at the start of a
 * group; NaN when none is valid), pinned by riptable/tests/test_groupbyops.py:901-1110 against np_rolling_nanquantile.
 * Windows up to 1024 rows. */
int rtb_groupby_rolling_quantile(const void* values, int vdtype, const void* ikey, int kdtype, const int32_t* igroup,
                                 const int32_t* ifirstgroup, const int32_t* ncountgroup, int64_t nrows,
                                 int64_t unique_rows, int64_t window, double q, double* out, void* stream);

/* ---- a10: rc.MakeiFirst(key, n, filter, mode) / rc.MakeiNext(key, n, mode) — riptable/rt_numpy.py:1791-1854 */
int rtb_make_ifirst(const void* ikey, int kdtype, int64_t nrows, int64_t n, const uint8_t* filter, int mode,
                    int32_t* out /* n+1 */, void* stream);
size_t rtb_make_inext_workspace_bytes(int64_t nrows, int64_t n);
int rtb_make_inext(const void* ikey, int kdtype, int64_t nrows, int64_t n, int mode, int32_t* out /* nrows */,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- a11: rc.LexSort32(list, cutoffs=, index=, ascending=) — riptable/rt_numpy.py:2670-2671, 2200-2202
 * Stable lexicographic argsort; LAST key is primary (numpy convention); NaN last.
 *   index     int32[nindex] rows to sort (NULL = all rows, nindex = nrows)
 *   ascending_host int8[nkeys] per key (NULL = all ascending)
 *   perm_out  int32[nindex] */
size_t rtb_lexsort_workspace_bytes(int64_t nrows, int64_t nindex, int nkeys, const int32_t* itemsizes_host);
int rtb_lexsort32(const rtb_column* keys, int nkeys, int64_t nrows, const int32_t* index, int64_t nindex,
                  const int8_t* ascending_host, int32_t* perm_out, void* workspace, size_t workspace_bytes,
                  void* stream);

/* rc.LexSort64 (riptable/rt_numpy.py:2199-2200, 2669-2670: the entry riptable calls above 2e9 rows): the same sort with
 * an int64 result and an int64 `index`; up to 2^32 - 2 rows (row ids are uint32 inside the radix engine). */
size_t rtb_lexsort64_workspace_bytes(int64_t nrows, int64_t nindex, int nkeys, const int32_t* itemsizes_host);
int rtb_lexsort64(const rtb_column* keys, int nkeys, int64_t nrows, const int64_t* index, int64_t nindex,
                  const int8_t* ascending_host, int64_t* perm_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- multi-GPU sample sort: dest_out[r] = number of the `nsplit` splitter rows that sort at or before
 * row r under LexSort32's order (last key primary, NaN last, per-key ascending flags) with the global row id
 * (row_offset + r vs splitter_rows[j], a device array) as the final tie-break. */
int rtb_lex_partition(const rtb_column* keys, int nkeys, int64_t nrows, const int8_t* ascending_host,
                      const rtb_column* splitters, int nsplit, const int64_t* splitter_rows, int64_t row_offset,
                      int32_t* dest_out, void* stream);

/* ---- a12: rc.GroupFromLexSort(iGroup, value_array, cutoffs=, base_index=) — riptable/rt_numpy.py:2207
 *   ikey_out int32[nrows] (rows not in igroup get 0), ifirstkey_out int32[nindex] (first U valid),
 *   ncountgroup_out int32[nindex+1] (first U+1 valid, slot 0 = 0). */
size_t rtb_group_from_lexsort_workspace_bytes(int64_t nrows, int64_t nindex);
int rtb_group_from_lexsort(const int32_t* igroup, int64_t nindex, const rtb_column* keys, int nkeys, int64_t nrows,
                           int base_index, int32_t* ikey_out, int32_t* ifirstkey_out, int32_t* ncountgroup_out,
                           int64_t* unique_count_host, void* workspace, size_t workspace_bytes, void* stream);

/* ---- `cutoffs=` of rc.LexSort32 / rc.GroupFromLexSort / rc.GroupByPack32 — riptable/rt_numpy.py:2199-2207, 1972.
 * cutoffs (int64 cumulative row ends; a DEVICE array here) splits the rows into contiguous partitions that behave as
 * arrays of their own (the meaning riptable/tests/test_pgroupby.py:7-86 pins for the hash case; the sort / pack cases
 * are unpinned by the reference's tests and derived from it).  The shim runs ONE sort / pack over all rows with the
 * partition id as an additional most significant key and uses these helpers to translate between global and
 * partition-relative numbers:
 *   rtb_partition_ids        pid_out[r] = partition of row r
 *   rtb_partition_shift      values[i] -/+= first row of the partition POSITION i lies in (to_relative = 1 / 0)
 *   rtb_partition_count_rows count_out[p] = number of entries of rows[] lying in partition p
 *   rtb_partition_localize   iKey numbered over all partitions (partition-major, base 1) -> numbering that restarts in
 *                            every partition, in base `base_index` (bin_base[p] = bins before partition p);
 *                            ifirstkey (may be NULL) -> partition-relative rows
 *   rtb_partition_max_ikey   umax_out[p] = largest iKey in partition p; *bad_out counts negative keys
 *   rtb_partition_compose_ikey out[r] = bin_off[partition of r] + iKey[r] */
int rtb_partition_ids(int64_t n, const int64_t* cutoffs, int64_t ncutoffs, int32_t* pid_out, void* stream);
int rtb_partition_shift(int32_t* values, int64_t n, const int32_t* pid, const int64_t* cutoffs, int64_t ncutoffs,
                        int to_relative, void* stream);
int rtb_partition_count_rows(const int32_t* rows, int64_t m, const int32_t* pid, int64_t ncutoffs, int32_t* count_out,
                             void* stream);
int rtb_partition_localize(int32_t* ikey, int64_t n, int32_t* ifirstkey, int64_t nbins, const int32_t* pid,
                           const int32_t* bin_base, const int64_t* cutoffs, int64_t ncutoffs, int base_index, void* stream);
int rtb_partition_max_ikey(const void* ikey, int kdtype, int64_t n, const int32_t* pid, int64_t ncutoffs, int32_t* umax_out,
                           uint32_t* bad_out, void* stream);
int rtb_partition_compose_ikey(const void* ikey, int kdtype, int64_t n, const int32_t* pid, const int64_t* bin_off,
                               int32_t* out, void* stream);

/* ---- 8f.4: join-index construction — riptable/rt_merge.py:734-1190 (`_build_left_fancyindex`,
 * `_build_right_fancyindex`) and its numba loops :611-665, 668-731, 887-975; no rc entry exists for these (riptable
 * runs them in Python / numba), so riptable_b200/merge.py mirrors `rt.merge_indices` and calls:
 *   rtb_valid_mask        out[r] = key r is usable: not NaN / not the integer sentinel (`_create_column_valid_mask`,
 *                         rt_merge.py:1848-1872); mode 0 sets, 1 ANDs into, 2 ORs into `out`
 *   rtb_merge_row_counts  per left row: keyid_out = the 1-based right key id its key maps to through `keymap`
 *                         (int32[nmap] indexed by left key id - 1; negative = no such key on the right) or the int32
 *                         invalid; count_out = rows it contributes: that right group's size, or nonmatched_count
 *   rtb_merge_scan_counts exclusive scan of the counts and their total (host); RTB_ERR_ARG when the total does not fit an
 *                         int32 row index (the total is still reported)
 *   rtb_merge_expand      left_out[off .. off + count) = the left row (left_rows[i], or i when NULL);
 *                         right_out[...] = right_igroup[right_ifirstgroup[keyid] ...], or the int32 invalid for an
 *                         unmatched row.  Either output may be NULL.
 * rtb_merge_row_counts: a row whose key id is 0 (an invalid key) contributes invalid_row_count rows, a row whose key has no
 * partner nonmatched_count rows (0 for an inner join, 1 for a left join). */
int rtb_valid_mask(const void* col, int dtype, int64_t n, int mode, uint8_t* out, void* stream);
int rtb_merge_row_counts(const int32_t* left_ikey, int64_t n, const int32_t* keymap, int64_t nmap,
                         const int32_t* right_ncountgroup, int64_t ngroups, int64_t nonmatched_count,
                         int64_t invalid_row_count, int32_t* keyid_out, uint32_t* count_out, void* stream);
size_t rtb_merge_scan_workspace_bytes(int64_t n);
int rtb_merge_scan_counts(const uint32_t* counts, int64_t n, uint32_t* offsets_out, int64_t* total_host, void* workspace,
                          size_t workspace_bytes, void* stream);
int rtb_merge_expand(const int32_t* keyid, const uint32_t* offsets, const uint32_t* counts, int64_t n,
                     const int32_t* left_rows, const int32_t* right_ifirstgroup, const int32_t* right_igroup,
                     int32_t* left_out, int32_t* right_out, void* stream);

/* ---- a13: rc.ReverseShuffle(perm) — riptable/rt_grouping.py:1659.  out[perm[i]] = i. */
int rtb_reverse_shuffle(const int32_t* perm, int64_t n, int32_t* out, void* stream);

/* ---- multi-GPU helper: out[i] = map[ikey[i]] (rc.ReIndexGroups semantics, riptable/rt_grouping.py:378-407) */
int rtb_remap_ikey(const int32_t* ikey, int64_t nrows, const int32_t* map, int64_t nmap, int32_t* out, void* stream);

/* ---- 8f.3: rc.ReIndexGroups(ikey, uikey, u_cutoffs, i_cutoffs) — riptable/rt_grouping.py:378 (semantics spelled out by
 * the Python fallback at :380-407).  Partition p owns ikey rows [i_cutoffs[p-1], i_cutoffs[p]) and uikey entries
 * [u_cutoffs[p-1], u_cutoffs[p]); out[r] = uikey_slice_p[ikey[r] - 1], and 0 stays 0.  Rows past the last cutoff -> 0. */
int rtb_reindex_groups(const int32_t* ikey, int64_t nrows, const int32_t* uikey, int64_t nuikey,
                       const int64_t* u_cutoffs_host, const int64_t* i_cutoffs_host, int64_t ncutoffs, int32_t* out,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIPTABLE_B200_H */
