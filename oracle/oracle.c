/* CPU ORACLE in C (test infrastructure and the timed CPU baseline; NOT a product path).
 *
 * Restates the algorithm class riptable documents for this path:
 *   - rc.MultiKeyGroupBy32: "linear hashing algorithm ... bin each group according to first
 *     appearance" (riptable/rt_numpy.py:1983-1986) — a sequential linear-probe table; a key's bin is
 *     the count of distinct keys seen when it first appears.  Without `cutoffs` the reference runs
 *     this stage on one thread (cutoffs is its "parallel processing" handle, rt_numpy.py:1914-1915).
 *   - rc.GroupByAll32 sum: threads own disjoint bin ranges and each scans all rows — the structure of
 *     riptable's in-tree mirror (riptable/rt_groupbynumba.py:30-50 build_core_list, :139-160 _numbasum).
 * riptide_cpp's own source is not in /root/reference (pinned >=1.19.0,<2, pyproject.toml:15), so this
 * is a port ("kind": "port" in bench.py), checked against the NumPy oracle in tests/test_c_oracle.py,
 * which in turn is pinned to the reference's golden vectors.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t mix(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

/* keys: n 64-bit key patterns (bytewise equality). filter may be NULL.
 * ikey_out int32[n] (base 1, 0 = filtered), ifirst_out int32[n] (first U valid). Returns U, or -1. */
int64_t orc_groupby_u64(const uint64_t* keys, const uint8_t* filter, int64_t n, int32_t* ikey_out,
                        int32_t* ifirst_out) {
  uint64_t cap = 1 << 16;
  uint64_t* tk = (uint64_t*)malloc(cap * sizeof(uint64_t));
  int32_t* tb = (int32_t*)calloc(cap, sizeof(int32_t)); /* 0 = empty, else bin */
  if (!tk || !tb) return -1;
  int64_t U = 0;
  for (int64_t r = 0; r < n; r++) {
    if (filter && !filter[r]) { ikey_out[r] = 0; continue; }
    if ((uint64_t)(U + 1) * 2 > cap) { /* grow and rehash */
      uint64_t ncap = cap * 4;
      uint64_t* nk = (uint64_t*)malloc(ncap * sizeof(uint64_t));
      int32_t* nb = (int32_t*)calloc(ncap, sizeof(int32_t));
      if (!nk || !nb) return -1;
      for (uint64_t i = 0; i < cap; i++)
        if (tb[i]) {
          uint64_t j = mix(tk[i]) & (ncap - 1);
          while (nb[j]) j = (j + 1) & (ncap - 1);
          nk[j] = tk[i]; nb[j] = tb[i];
        }
      free(tk); free(tb); tk = nk; tb = nb; cap = ncap;
    }
    uint64_t k = keys[r], j = mix(k) & (cap - 1);
    for (;;) {
      if (!tb[j]) { tk[j] = k; tb[j] = (int32_t)(++U); ifirst_out[U - 1] = (int32_t)r; ikey_out[r] = (int32_t)U; break; }
      if (tk[j] == k) { ikey_out[r] = tb[j]; break; }
      j = (j + 1) & (cap - 1);
    }
  }
  free(tk); free(tb);
  return U;
}

/* The same grouping with riptable's own parallel handle: contiguous row partitions grouped independently
 * (`cutoffs`, riptable/rt_numpy.py:1914-1915 "cutoff array for parallel processing") and merged the way
 * hstack_groupings does (riptable/rt_grouping.py:277-413): re-group the stacked per-partition uniques in partition
 * order, then re-index every partition's iKey.  Result identical to orc_groupby_u64. */
int64_t orc_groupby_u64_parallel(const uint64_t* keys, const uint8_t* filter, int64_t n, int32_t* ikey_out,
                                 int32_t* ifirst_out, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (n < (int64_t)threads * 1024) return orc_groupby_u64(keys, filter, n, ikey_out, ifirst_out);
  int64_t* lo = (int64_t*)malloc(sizeof(int64_t) * (threads + 1));
  int64_t* uc = (int64_t*)calloc(threads, sizeof(int64_t));
  int32_t** lfirst = (int32_t**)calloc(threads, sizeof(int32_t*));
  int32_t** map = (int32_t**)calloc(threads, sizeof(int32_t*));
  if (!lo || !uc || !lfirst || !map) return -1;
  for (int t = 0; t <= threads; t++) lo[t] = n * t / threads;
  int failed = 0;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
  for (int t = 0; t < threads; t++) {
    int64_t m = lo[t + 1] - lo[t];
    lfirst[t] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    if (!lfirst[t]) { failed = 1; continue; }
    uc[t] = orc_groupby_u64(keys + lo[t], filter ? filter + lo[t] : NULL, m, ikey_out + lo[t], lfirst[t]);
    if (uc[t] < 0) failed = 1;
  }
  if (failed) return -1;
  /* merge: stacked uniques in partition order -> global first-appearance bins */
  uint64_t cap = 1 << 16;
  int64_t total = 0;
  for (int t = 0; t < threads; t++) total += uc[t];
  while (cap < (uint64_t)total * 2 + 2) cap <<= 1;
  uint64_t* tk = (uint64_t*)malloc(cap * sizeof(uint64_t));
  int32_t* tb = (int32_t*)calloc(cap, sizeof(int32_t));
  if (!tk || !tb) return -1;
  int64_t U = 0;
  for (int t = 0; t < threads; t++) {
    map[t] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(uc[t] + 1));
    if (!map[t]) return -1;
    map[t][0] = 0;
    for (int64_t j = 0; j < uc[t]; j++) {
      int64_t row = lo[t] + lfirst[t][j];
      uint64_t k = keys[row], i = mix(k) & (cap - 1);
      for (;;) {
        if (!tb[i]) { tk[i] = k; tb[i] = (int32_t)(++U); ifirst_out[U - 1] = (int32_t)row; map[t][j + 1] = (int32_t)U; break; }
        if (tk[i] == k) { map[t][j + 1] = tb[i]; break; }
        i = (i + 1) & (cap - 1);
      }
    }
  }
#pragma omp parallel for num_threads(threads) schedule(static, 1)
  for (int t = 1; t < threads; t++) { /* partition 0 is already in global numbering */
    const int32_t* m = map[t];
    for (int64_t r = lo[t]; r < lo[t + 1]; r++) ikey_out[r] = m[ikey_out[r]];
  }
  for (int t = 0; t < threads; t++) { free(lfirst[t]); free(map[t]); }
  free(lfirst); free(map); free(lo); free(uc); free(tk); free(tb);
  return U;
}

/* sum of f64 values per bin; out[nb]; bins outside [lo,hi) left 0.  Threads own bin ranges. */
void orc_sum_f64(const double* v, const int32_t* ikey, int64_t n, int64_t nb, int64_t lo, int64_t hi, double* out,
                 int threads) {
  memset(out, 0, sizeof(double) * (size_t)nb);
  if (threads < 1) threads = 1;
  if (threads > hi - lo) threads = (int)(hi - lo > 0 ? hi - lo : 1);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
  for (int t = 0; t < threads; t++) {
    int64_t span = hi - lo, q = span / threads, rem = span % threads;
    int64_t blo = lo + t * q + (t < rem ? t : rem), bhi = blo + q + (t < rem ? 1 : 0);
    for (int64_t r = 0; r < n; r++) {
      int64_t b = ikey[r];
      if (b >= blo && b < bhi) out[b] += v[r];
    }
  }
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
