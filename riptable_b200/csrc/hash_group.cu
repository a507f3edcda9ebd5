// a1: MultiKeyGroupBy32 — hash grouping with first-appearance bin numbering, on sm_100a.
//
// Replaces rc.MultiKeyGroupBy32 (riptable/rt_numpy.py:2073-2075).  Design:
//
//   * Open-addressing table of 16-byte slots {key|tag, minrow, bin}, probed linearly from an EVEN slot index, so
//     the table is a sequence of 32-byte (one L2 sector) buckets of two slots.  The single-key steady-state
//     probe reads a whole bucket with one sm_100a 256-bit load (LDG.256): measured 286 G random 32 B probes/s
//     against 226 G/s for 16 B loads (profiles/microbench/l2_random2.cu), and two slots per probe cut the
//     expected probes per row from 1.16 to 1.03 at 24 % load.
//   * PROGRESSIVE ROUNDS over geometrically growing row ranges [P_k, P_{k+1}).  A key already
//     finalised in an earlier round resolves with one probe and iKey is written immediately, so the
//     steady state moves key_bytes + 4 bytes per row — the algorithmic minimum.  Keys first seen in a
//     round are "provisional": the row stores -(slot+1), the slot tracks min(row) with atomicMin.
//   * After a round, provisional slots are ranked by min row through a bitmap over the round's rows
//     and a device-wide popcount scan; that rank IS first-appearance order because every key of an
//     earlier round has a smaller first row than any key new to this round.  A fix-up pass rewrites
//     the provisional rows of that round only.
//   * The table is resized between rounds to ~2x the projected unique count so that it stays
//     L2-resident for low/medium cardinalities; a round that runs out of room aborts and is redone
//     after growing (bounded by the geometric schedule).
//
// Two key layouts:
//   direct : one key column of 1/2/4/8 bytes, the key itself (zero-extended) lives in the slot.
//   byref  : anything else (multi-key, strings). The slot holds (hash_hi32 << 32 | representative row)
//            and equality is verified bytewise against the representative row's key columns.
#include <math.h>
#include <stdlib.h>

#include <cooperative_groups.h>
#include <type_traits>

#include "common.cuh"
#include "keycols.cuh"
#include "reduce_internal.cuh"
#include "values.cuh"

namespace rtb {

struct __align__(16) Slot {
  unsigned long long key;  // direct: key bits; byref: (hash_hi32 << 32) | rep_row.  ~0 = empty
  uint32_t minrow;         // smallest row seen for this key (0xFFFFFFFF until the first atomicMin lands)
  uint32_t nbin;           // ~bin: 0xFFFFFFFF while provisional, ~(final 1-based bin) afterwards (memset 0xFF = fresh)
  __device__ __forceinline__ uint32_t bin() const { return ~nbin; }
};
static_assert(sizeof(Slot) == 16, "slot must be one 128-bit load");

constexpr unsigned long long kEmpty = ~0ull;
struct GbCounters {
  uint32_t newcount;  // slots claimed in this round
  uint32_t abort;     // set when the round ran out of table room
  uint32_t total;     // scratch for scan totals
  uint32_t pad;
};

// ---- hashing ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_u64(unsigned long long k) {
  uint32_t lo = (uint32_t)k, hi = (uint32_t)(k >> 32);
  uint32_t a = lo * 0x9E3779B1u;
  uint32_t b = hi * 0x85EBCA77u;
  uint32_t h = a ^ __funnelshift_l(b, b, 16);
  h ^= h >> 15;
  h *= 0xC2B2AE3Du;
  h ^= h >> 13;
  return h;
}

__device__ __forceinline__ Slot load_slot(const Slot* s) {
  // Plain (L1-allocating) 128-bit read.  A stale L1 copy is harmless: keys of finalised slots never change within a
  // launch, a stale "empty" is corrected by the atomicCAS that follows, and a stale minrow is only ever
  // too large, which at worst issues a redundant atomicMin.
  uint4 w;
  asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(s) : "memory");
  Slot r;
  r.key = ((unsigned long long)w.y << 32) | w.x;
  r.minrow = w.z;
  r.nbin = w.w;
  return r;
}

// One bucket = two slots = one 32-byte sector, fetched with a single 256-bit load.
struct Bucket {
  unsigned long long k0, w0, k1, w1;  // w = minrow | (nbin << 32)
};
// RTB_TABLE_EL=1 (measurement build): the bucket load carries an L2 evict-last policy.
#ifndef RTB_TABLE_EL
#define RTB_TABLE_EL 0
#endif
__device__ __forceinline__ Bucket load_bucket(const Slot* table, uint32_t b) {
  Bucket p;
#if RTB_TABLE_EL
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.L2::cache_hint.v4.b64 {%0,%1,%2,%3}, [%4], %5;"
               : "=l"(p.k0), "=l"(p.w0), "=l"(p.k1), "=l"(p.w1) : "l"(table + 2 * (size_t)b), "l"(pol));
#else
  asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];"
               : "=l"(p.k0), "=l"(p.w0), "=l"(p.k1), "=l"(p.w1) : "l"(table + 2 * (size_t)b));
#endif
  return p;
}

// A row landed on a provisional slot (or just claimed it): track the first row, return -(slot+1).
__device__ __forceinline__ int32_t provisional(Slot* table, uint32_t idx, uint32_t row, uint32_t seen_minrow) {
  if (row < seen_minrow) atomicMin(&table[idx].minrow, row);
  return -(int32_t)(idx + 1);
}

// Counter bump aggregated over the threads of the warp that arrive here together: one same-address atomic per
// group instead of one per thread (a discovery round over 20 M new keys spent 27 ms serialised on this counter).
__device__ __forceinline__ uint32_t aggregated_inc(uint32_t* counter) {
  cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
  uint32_t base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(counter, (uint32_t)g.size());
  return g.shfl(base, 0) + g.thread_rank();
}

__device__ __forceinline__ bool claim_new(GbCounters* ctr, uint32_t* newlist, uint32_t idx, uint32_t new_limit) {
  uint32_t pos = aggregated_inc(&ctr->newcount);
  if (pos >= new_limit) {
    ctr->abort = 1;
    return false;
  }
  newlist[pos] = idx;
  return true;
}

constexpr uint32_t kNoClaim = 0xFFFFFFFFu;

// Hands the slots its caller's threads claimed (kNoClaim = none) to the round's new-key list: one bump of the shared
// counter for all the threads that arrive together and all N of their rows.  (The counter is a single address; bumped once per coalesced
// group *per row* it serialised a 1 Mi-row first round with 650 K new keys to 138 us.)
template <int N>
__device__ __forceinline__ void claim_new_n(GbCounters* ctr, uint32_t* newlist, const uint32_t (&claimed)[N], uint32_t new_limit) {
  cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
  uint32_t b[N];
  uint32_t total = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    b[i] = g.ballot(claimed[i] != kNoClaim);
    total += __popc(b[i]);
  }
  if (total == 0) return;
  uint32_t base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(&ctr->newcount, total);
  base = g.shfl(base, 0);
  const uint32_t below = (1u << g.thread_rank()) - 1u;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (claimed[i] != kNoClaim) {
      const uint32_t pos = base + __popc(b[i] & below);
      if (pos < new_limit) newlist[pos] = claimed[i];
      else ctr->abort = 1;  // out of room: the round is redone with a larger table, this entry is dropped by the re-stage
    }
    base += __popc(b[i]);
  }
}

// ---- direct path -----------------------------------------------------------------------------
// General resolution of one key, slot by slot from slot `idx` (even): handles matches, provisional slots,
// insertion of new keys and the key that equals the empty marker.  `mask` = capacity - 1 (capacity in slots).
__device__ __noinline__ int32_t resolve_direct(Slot* table, uint32_t mask, unsigned long long key, uint32_t idx,
                                               uint32_t row, uint32_t* newlist, GbCounters* ctr, uint32_t new_limit) {
  if (key == kEmpty) {
    // the one key that collides with the empty marker lives in the extra slot table[cap]
    uint32_t sp = mask + 1;
    Slot* ss = &table[sp];
    Slot cur = load_slot(ss);
    if (cur.bin()) return (int32_t)cur.bin();
    if (cur.key == kEmpty) {
      unsigned long long old = atomicCAS(&ss->key, kEmpty, 0ull);
      if (old == kEmpty && !claim_new(ctr, newlist, sp, new_limit)) return 0;
    }
    return provisional(table, sp, row, cur.minrow);
  }
  for (uint32_t probes = 0; probes <= mask; probes++) {
    Slot s = load_slot(&table[idx]);
    if (s.key == key) {
      if (s.bin()) return (int32_t)s.bin();
      return provisional(table, idx, row, s.minrow);
    }
    if (s.key == kEmpty) {
      // once the round is out of room (it will be redone) nothing more is inserted, so the table cannot fill up
      if (*(volatile uint32_t*)&ctr->abort) return 0;
      unsigned long long old = atomicCAS(&table[idx].key, kEmpty, key);
      if (old == kEmpty) {
        if (!claim_new(ctr, newlist, idx, new_limit)) return 0;
        return provisional(table, idx, row, 0xFFFFFFFFu);
      }
      if (old == key) return provisional(table, idx, row, 0xFFFFFFFFu);
    }
    idx = (idx + 1) & mask;
  }
  ctr->abort = 1;  // table full (cannot happen while new_limit < capacity)
  return 0;
}

// Steady-state classification of one fetched bucket: > 0 final bin, 0 = both slots hold other keys (go on to the
// next bucket), -1 = needs the general path (an empty or provisional slot, or the empty-marker key).
__device__ __forceinline__ int32_t classify(const Bucket& p, unsigned long long key) {
  if (p.k0 == key) {
    uint32_t bin = ~(uint32_t)(p.w0 >> 32);
    return bin ? (int32_t)bin : -1;
  }
  if (p.k1 == key) {
    uint32_t bin = ~(uint32_t)(p.w1 >> 32);
    return bin ? (int32_t)bin : -1;
  }
  return (p.k0 == kEmpty || p.k1 == kEmpty) ? -1 : 0;
}

template <int ISZ>
__device__ __forceinline__ unsigned long long load_key1(const void* keys, int64_t r) {
  if (ISZ == 8) return ((const unsigned long long*)keys)[r];
  if (ISZ == 4) return ((const uint32_t*)keys)[r];
  if (ISZ == 2) return ((const uint16_t*)keys)[r];
  return ((const uint8_t*)keys)[r];
}

// A warp owns 128-row tiles; every lane holds 4 keys.  Row of lane `l`, key j inside a tile:
//   PAIR  (8-byte keys, or a fused 8-byte value column): vector loads that each cover contiguous memory per warp
//                                                        -> rows 2l, 2l+1, 64+2l, 65+2l
//   else  : one vector load of 4 consecutive rows        -> rows 4l .. 4l+3
template <int ISZ, bool PAIR = (ISZ == 8)>
__device__ __forceinline__ int tile_row(int lane, int j) {
  if (PAIR) return (j < 2 ? 2 * lane + j : 64 + 2 * lane + (j - 2));
  return 4 * lane + j;
}
template <int ISZ, bool PAIR = (ISZ == 8)>
__device__ __forceinline__ void load_tile_keys(const void* keys, int64_t tb, int lane, unsigned long long k[4]) {
  if (ISZ == 8) {
    uint4 a = ld_stream_u4((const uint8_t*)keys + (tb + 2 * lane) * 8);
    uint4 b = ld_stream_u4((const uint8_t*)keys + (tb + 64 + 2 * lane) * 8);
    k[0] = ((unsigned long long)a.y << 32) | a.x; k[1] = ((unsigned long long)a.w << 32) | a.z;
    k[2] = ((unsigned long long)b.y << 32) | b.x; k[3] = ((unsigned long long)b.w << 32) | b.z;
  } else if (PAIR) {
    if (ISZ == 4) {
      uint2 a = ld_stream_u2((const uint8_t*)keys + (tb + 2 * lane) * 4);
      uint2 b = ld_stream_u2((const uint8_t*)keys + (tb + 64 + 2 * lane) * 4);
      k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y;
    } else if (ISZ == 2) {
      uint32_t a = ld_stream_u32((const uint8_t*)keys + (tb + 2 * lane) * 2);
      uint32_t b = ld_stream_u32((const uint8_t*)keys + (tb + 64 + 2 * lane) * 2);
      k[0] = a & 0xffff; k[1] = a >> 16; k[2] = b & 0xffff; k[3] = b >> 16;
    } else {
      uint32_t a = *(const uint16_t*)((const uint8_t*)keys + tb + 2 * lane);
      uint32_t b = *(const uint16_t*)((const uint8_t*)keys + tb + 64 + 2 * lane);
      k[0] = a & 0xff; k[1] = a >> 8; k[2] = b & 0xff; k[3] = b >> 8;
    }
  } else if (ISZ == 4) {
    uint4 a = ld_stream_u4((const uint8_t*)keys + (tb + 4 * lane) * 4);
    k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w;
  } else if (ISZ == 2) {
    uint2 a = ld_stream_u2((const uint8_t*)keys + (tb + 4 * lane) * 2);
    k[0] = a.x & 0xffff; k[1] = a.x >> 16; k[2] = a.y & 0xffff; k[3] = a.y >> 16;
  } else {
    uint32_t a = ld_stream_u32((const uint8_t*)keys + tb + 4 * lane);
    k[0] = a & 0xff; k[1] = (a >> 8) & 0xff; k[2] = (a >> 16) & 0xff; k[3] = a >> 24;
  }
}

// ---- fused per-bin sum (rtb_multikey_groupby_sum32) ---------------------------------------------------------
// The steady-state round can add every row's value to its bin in the same pass that writes iKey: the probe already
// has the final bin in a register, so the separate reduction pass (4 B/row of iKey re-read + the value stream + a
// second kernel's worth of L2 round trips) disappears.  Rows that land on a provisional slot are added by gb_fixup.
struct GbFuse {
  const void* values;        // one value per row
  unsigned long long* acc;   // per-bin accumulators (f64 bits, or 64-bit wrap-around integers), zeroed as bins are created
  unsigned long long* snap;  // as large as acc: the accumulators as they stood before the current fused round
  int vdt;                   // rtb_dtype of the values
  int vsz;                   // bytes per value
  int nanskip;               // GB_NANSUM: NaN / integer sentinels are skipped
  int isfloat;               // accumulate in f64 (float columns), else in int64
  int bin_low;               // 1: filtered rows are skipped; 0: they are summed into bin 0 (riptable's showfilter)
};
__device__ __forceinline__ void fuse_add(const GbFuse& f, int32_t bin, unsigned long long raw) {
  if (f.nanskip && raw_invalid(raw, f.vdt)) return;
  if (f.isfloat) atomicAdd((double*)&f.acc[bin], raw_to_f64(raw, f.vdt));
  else atomicAdd(&f.acc[bin], (unsigned long long)raw_to_i64(raw, f.vdt));
}

// One round of the direct path over rows [lo, hi).  lo is a multiple of 128; the key / filter / ikey base
// pointers are 16-byte aligned (checked on the host).
// FUSE: 0 = grouping only; 1 / 2 = also sum an 8-byte / 4-byte value column into `fz.acc` (see GbFuse).
// Occupancy is what this kernel lives on: the grouping-only form is compiled for 4 CTAs per SM (64 registers; the same
// code left at 82 registers / 3 CTAs ran its rounds at 142 instead of 202 G rows/s).  The fused form carries four
// 64-bit values more per thread: with the next tile's keys prefetched it needs 78 registers (3 CTAs per SM), without
// the prefetch it fits 64 registers (4 CTAs per SM) -- both are built, the plan picks (GbPlan::fuse_prefetch).
template <int ISZ, bool PREFETCH, int FUSE>
__global__ void __launch_bounds__(256, (FUSE && PREFETCH) ? 3 : 4) gb_round_direct(const void* __restrict__ keys, const uint8_t* __restrict__ filter,
                                                       int32_t* __restrict__ ikey, Slot* table, uint32_t mask,
                                                       int64_t lo, int64_t hi, uint32_t* newlist, GbCounters* ctr,
                                                       uint32_t new_limit, GbFuse fz) {
  constexpr bool PAIR = ISZ == 8 || FUSE == 1;
  constexpr int VSZ = FUSE == 1 ? 8 : 4;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t bmask = mask >> 1;  // buckets - 1
  const int64_t step = nwarps * 128;
  int64_t tb = lo + gw * 128;
  unsigned long long k[4], kn[4] = {0, 0, 0, 0};
  if (PREFETCH && tb + 128 <= hi) load_tile_keys<ISZ, PAIR>(keys, tb, lane, k);
  for (; tb + 128 <= hi; tb += step) {
    if (PREFETCH) {
      if (tb + step + 128 <= hi) load_tile_keys<ISZ, PAIR>(keys, tb + step, lane, kn);  // in flight during the probes
    } else {
      load_tile_keys<ISZ, PAIR>(keys, tb, lane, k);
    }
    uint32_t fbits = 0xF;
    if (filter) {
      if (PAIR) {
        uint32_t fa = *(const uint16_t*)(filter + tb + 2 * lane), fb = *(const uint16_t*)(filter + tb + 64 + 2 * lane);
        fbits = ((fa & 0xff) ? 1u : 0u) | ((fa >> 8) ? 2u : 0u) | ((fb & 0xff) ? 4u : 0u) | ((fb >> 8) ? 8u : 0u);
      } else {
        uint32_t f = ld_stream_u32(filter + tb + 4 * lane);
        fbits = ((f & 0xff) ? 1u : 0u) | ((f & 0xff00) ? 2u : 0u) | ((f & 0xff0000) ? 4u : 0u) | ((f >> 24) ? 8u : 0u);
      }
    }
    uint32_t b[4];
    Bucket p[4];
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = hash_u64(k[i]) & bmask;
#pragma unroll
    for (int i = 0; i < 4; i++) p[i] = load_bucket(table, b[i]);
    unsigned long long v[4] = {0, 0, 0, 0};
    if (FUSE) {  // the value stream rides behind the probes
      if (PAIR) {
        load_raw2(fz.values, VSZ, tb + 2 * lane, v);
        load_raw2(fz.values, VSZ, tb + 64 + 2 * lane, v + 2);
      } else {
        load_raw4(fz.values, VSZ, tb + 4 * lane, v);
      }
    }
    int32_t o[4];
    uint32_t pend = 0;
    // branch-free common case: the key sits in its home bucket with a final bin.  (The empty-marker key can
    // only "match" an empty slot, whose bin field reads 0, so it falls through to the general section.)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool m0 = p[i].k0 == k[i], m1 = p[i].k1 == k[i];
      const unsigned long long w = m0 ? p[i].w0 : p[i].w1;
      const uint32_t bin = ~(uint32_t)(w >> 32);
      const bool ok = (m0 | m1) & (bin != 0u);
      o[i] = ok ? (int32_t)bin : 0;
      if (!ok) pend |= 1u << i;
    }
    pend &= fbits;
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (!((fbits >> i) & 1u)) o[i] = 0;
    if (pend) {
      // general section: next buckets for keys whose home bucket is full of other keys, the insert path for the rest
      uint32_t slow = 0;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if ((pend >> i) & 1u) {
          int32_t c = k[i] == kEmpty ? -1 : classify(p[i], k[i]);
          if (c < 0) {
            slow |= 1u << i;
            pend &= ~(1u << i);
          }
        }
      uint32_t hops = 0;
      while (pend) {
        if (++hops > bmask) {  // cannot happen while the table has free slots; never spin
          ctr->abort = 1;
          break;
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
          if ((pend >> i) & 1u) {
            b[i] = (b[i] + 1) & bmask;
            p[i] = load_bucket(table, b[i]);
          }
#pragma unroll
        for (int i = 0; i < 4; i++)
          if ((pend >> i) & 1u) {
            int32_t c = classify(p[i], k[i]);
            if (c > 0) {
              o[i] = c;
              pend &= ~(1u << i);
            } else if (c < 0) {
              slow |= 1u << i;
              pend &= ~(1u << i);
            }
          }
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
        if ((slow >> i) & 1u)
          o[i] = resolve_direct(table, mask, k[i], 2 * b[i], (uint32_t)(tb + tile_row<ISZ, PAIR>(lane, i)), newlist, ctr, new_limit);
    }
    if (PAIR) {
#if RTB_STREAM_CS
      st_stream_u2(ikey + tb + 2 * lane, make_uint2((uint32_t)o[0], (uint32_t)o[1]));
      st_stream_u2(ikey + tb + 64 + 2 * lane, make_uint2((uint32_t)o[2], (uint32_t)o[3]));
#else
      *(int2*)(ikey + tb + 2 * lane) = make_int2(o[0], o[1]);
      *(int2*)(ikey + tb + 64 + 2 * lane) = make_int2(o[2], o[3]);
#endif
    } else {
      st_stream_u4(ikey + tb + 4 * lane, make_uint4((uint32_t)o[0], (uint32_t)o[1], (uint32_t)o[2], (uint32_t)o[3]));
    }
    if (FUSE) {
      // final bins are summed here; provisional rows (o < 0) are summed by gb_fixup once their bin is known
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (o[i] >= fz.bin_low) fuse_add(fz, o[i], v[i]);
    }
    if (PREFETCH) {
#pragma unroll
      for (int i = 0; i < 4; i++) k[i] = kn[i];
    }
  }
  // the warp whose next tile is the partial one finishes the rows that are left
  if (tb < hi) {
    for (int64_t q = tb + lane; q < hi; q += 32) {
      int32_t o = 0;
      if (!filter || filter[q]) {
        unsigned long long kk = load_key1<ISZ>(keys, q);
        o = resolve_direct(table, mask, kk, 2 * (hash_u64(kk) & bmask), (uint32_t)q, newlist, ctr, new_limit);
      }
      ikey[q] = o;
      if (FUSE && o >= fz.bin_low) fuse_add(fz, o, load_raw1(fz.values, VSZ, q));
    }
  }
}

// ---- direct path, low cardinality: the whole table in shared memory ---------------------------------------
// When the table has at most kSmemCapMax slots every CTA copies the finalised entries into shared memory once and
// probes there (shared memory serves several divergent lookups per clock; a global probe costs one L1TEX
// wavefront per lane), so the round is bound by the key / iKey streams alone.  Entries that are not final in the
// snapshot (none in steady state) and keys missing from it fall through to the general global-memory path.
constexpr uint32_t kSmemCapMax = 16384;

template <int ISZ>
__global__ void __launch_bounds__(1024, 1) gb_round_smem(const void* __restrict__ keys, const uint8_t* __restrict__ filter,
                                                         int32_t* __restrict__ ikey, Slot* table, uint32_t mask,
                                                         int64_t lo, int64_t hi, uint32_t* newlist, GbCounters* ctr,
                                                         uint32_t new_limit) {
  typedef typename std::conditional<ISZ == 8, unsigned long long, uint32_t>::type KT;
  extern __shared__ __align__(16) unsigned char gb_smem[];
  const uint32_t cap = mask + 1;
  KT* skey = (KT*)gb_smem;
  uint32_t* sbin = (uint32_t*)(skey + cap);
  for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) {
    Slot e = table[i];
    skey[i] = (KT)e.key;
    sbin[i] = e.bin();  // 0 = empty or not final: a miss
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t bmask = mask >> 1;
  const int64_t step = nwarps * 128;
  int64_t tb = lo + gw * 128;
  unsigned long long k[4], kn[4] = {0, 0, 0, 0};
  if (tb + 128 <= hi) load_tile_keys<ISZ>(keys, tb, lane, k);
  for (; tb + 128 <= hi; tb += step) {
    if (tb + step + 128 <= hi) load_tile_keys<ISZ>(keys, tb + step, lane, kn);
    uint32_t fbits = 0xF;
    if (filter) {
      if (ISZ == 8) {
        uint32_t fa = *(const uint16_t*)(filter + tb + 2 * lane), fb = *(const uint16_t*)(filter + tb + 64 + 2 * lane);
        fbits = ((fa & 0xff) ? 1u : 0u) | ((fa >> 8) ? 2u : 0u) | ((fb & 0xff) ? 4u : 0u) | ((fb >> 8) ? 8u : 0u);
      } else {
        uint32_t f = ld_stream_u32(filter + tb + 4 * lane);
        fbits = ((f & 0xff) ? 1u : 0u) | ((f & 0xff00) ? 2u : 0u) | ((f & 0xff0000) ? 4u : 0u) | ((f >> 24) ? 8u : 0u);
      }
    }
    int32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      o[i] = 0;
      if ((fbits >> i) & 1u) {
        const uint32_t home = (hash_u64(k[i]) & bmask) * 2;
        uint32_t idx = home, found = 0;
        if (!(ISZ == 8 && k[i] == kEmpty)) {
          for (;;) {
            uint32_t b = sbin[idx];
            if (b == 0) break;
            if (skey[idx] == (KT)k[i]) {
              found = b;
              break;
            }
            idx = (idx + 1) & mask;
          }
        }
        o[i] = found ? (int32_t)found
                     : resolve_direct(table, mask, k[i], home, (uint32_t)(tb + tile_row<ISZ>(lane, i)), newlist, ctr, new_limit);
      }
    }
    if (ISZ == 8) {
      *(int2*)(ikey + tb + 2 * lane) = make_int2(o[0], o[1]);
      *(int2*)(ikey + tb + 64 + 2 * lane) = make_int2(o[2], o[3]);
    } else {
      st_stream_u4(ikey + tb + 4 * lane, make_uint4((uint32_t)o[0], (uint32_t)o[1], (uint32_t)o[2], (uint32_t)o[3]));
    }
#pragma unroll
    for (int i = 0; i < 4; i++) k[i] = kn[i];
  }
  if (tb < hi) {
    for (int64_t q = tb + lane; q < hi; q += 32) {
      int32_t o = 0;
      if (!filter || filter[q]) {
        unsigned long long kk = load_key1<ISZ>(keys, q);
        o = resolve_direct(table, mask, kk, 2 * (hash_u64(kk) & bmask), (uint32_t)q, newlist, ctr, new_limit);
      }
      ikey[q] = o;
    }
  }
}

// ---- direct path, discovery rounds ------------------------------------------------------------------
// Used while many keys are still new: 16-byte probes, everything inline (38 registers, full occupancy),
// the four first probes of a thread in flight together.
// Resolve one key.  `s` is the slot already fetched for idx (memory-level parallelism: the caller issues all its
// first probes before resolving any).  A slot this call claims for a new key is reported in `claimed` (kNoClaim
// otherwise); the caller hands it to the new-key list.
__device__ __forceinline__ int32_t resolve_insert(Slot* table, uint32_t mask, unsigned long long key, uint32_t idx,
                                                  Slot s, uint32_t row, GbCounters* ctr, uint32_t& claimed) {
  claimed = kNoClaim;
  if (key == kEmpty) {
    // the one key that collides with the empty marker lives in the extra slot table[cap]
    uint32_t sp = mask + 1;
    Slot* ss = &table[sp];
    Slot cur = load_slot(ss);
    if (cur.bin()) return (int32_t)cur.bin();
    if (cur.key == kEmpty) {
      unsigned long long old = atomicCAS(&ss->key, kEmpty, 0ull);
      if (old == kEmpty) claimed = sp;
    }
    return provisional(table, sp, row, cur.minrow);
  }
  for (uint32_t probes = 0; probes <= mask; probes++) {
    if (s.key == key) {
      if (s.bin()) return (int32_t)s.bin();
      return provisional(table, idx, row, s.minrow);
    }
    if (s.key == kEmpty) {
      unsigned long long old = atomicCAS(&table[idx].key, kEmpty, key);
      if (old == kEmpty) {
        claimed = idx;
        return provisional(table, idx, row, 0xFFFFFFFFu);
      }
      if (old == key) return provisional(table, idx, row, 0xFFFFFFFFu);
    }
    idx = (idx + 1) & mask;
    s = load_slot(&table[idx]);
  }
  ctr->abort = 1;  // table full (cannot happen while new_limit < capacity)
  return 0;
}

template <int ISZ>
__device__ __forceinline__ void load_keys4(const void* keys, int64_t r, unsigned long long k[4]) {
  if (ISZ == 8) {
    uint4 a = ld_stream_u4((const uint8_t*)keys + r * 8), b = ld_stream_u4((const uint8_t*)keys + r * 8 + 16);
    k[0] = ((unsigned long long)a.y << 32) | a.x; k[1] = ((unsigned long long)a.w << 32) | a.z;
    k[2] = ((unsigned long long)b.y << 32) | b.x; k[3] = ((unsigned long long)b.w << 32) | b.z;
  } else if (ISZ == 4) {
    uint4 a = ld_stream_u4((const uint8_t*)keys + r * 4);
    k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w;
  } else if (ISZ == 2) {
    uint2 a = ld_stream_u2((const uint8_t*)keys + r * 2);
    k[0] = a.x & 0xffff; k[1] = a.x >> 16; k[2] = a.y & 0xffff; k[3] = a.y >> 16;
  } else {
    uint32_t a = ld_stream_u32((const uint8_t*)keys + r);
    k[0] = a & 0xff; k[1] = (a >> 8) & 0xff; k[2] = (a >> 16) & 0xff; k[3] = a >> 24;
  }
}
// One round of the direct path over rows [lo, hi).  lo is a multiple of 4; the key / filter / ikey
// base pointers are 16-byte aligned for their 4-row vectors (checked on the host).
template <int ISZ>
__global__ void __launch_bounds__(256) gb_round_insert(const void* __restrict__ keys, const uint8_t* __restrict__ filter,
                                                       int32_t* __restrict__ ikey, Slot* table, uint32_t mask,
                                                       int64_t lo, int64_t hi, uint32_t* newlist, GbCounters* ctr,
                                                       uint32_t new_limit) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t r = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; r < hi; r += stride) {
    if (*(volatile uint32_t*)&ctr->abort) break;  // the round is out of room and will be redone: stop inserting
    if (r + 3 < hi) {
      unsigned long long k[4];
      load_keys4<ISZ>(keys, r, k);
      uint32_t f = filter ? ld_stream_u32(filter + r) : 0x01010101u;
      uint32_t idx[4];
      Slot s[4];
#pragma unroll
      for (int i = 0; i < 4; i++) idx[i] = (hash_u64(k[i]) & (mask >> 1)) * 2;
#pragma unroll
      for (int i = 0; i < 4; i++) s[i] = load_slot(&table[idx[i]]);
      int32_t o[4];
      uint32_t claimed[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        claimed[i] = kNoClaim;
        o[i] = ((f >> (8 * i)) & 0xff) ? resolve_insert(table, mask, k[i], idx[i], s[i], (uint32_t)(r + i), ctr, claimed[i]) : 0;
      }
      claim_new_n<4>(ctr, newlist, claimed, new_limit);
      st_stream_u4(ikey + r, make_uint4((uint32_t)o[0], (uint32_t)o[1], (uint32_t)o[2], (uint32_t)o[3]));
    } else {
      for (int64_t q = r; q < hi; q++) {
        int32_t o = 0;
        if (!filter || filter[q]) {
          unsigned long long k = load_key1<ISZ>(keys, q);
          uint32_t idx = (hash_u64(k) & (mask >> 1)) * 2;
          uint32_t claimed;
          o = resolve_insert(table, mask, k, idx, load_slot(&table[idx]), (uint32_t)q, ctr, claimed);
          if (claimed != kNoClaim && !claim_new(ctr, newlist, claimed, new_limit)) o = 0;
        }
        ikey[q] = o;
      }
    }
  }
}

// ---- byref path ------------------------------------------------------------------------------
// General resolution of one key row from slot `idx` on: tag compare, bytewise verify against the representative
// row, insert when an empty slot is reached.
__device__ __forceinline__ int32_t resolve_byref(const KeyCols& kc, Slot* table, uint32_t mask, int64_t r, uint32_t tag,
                                              uint32_t idx, uint32_t* newlist, GbCounters* ctr, uint32_t new_limit) {
  const unsigned long long mine = ((unsigned long long)tag << 32) | (uint32_t)r;
  for (uint32_t probes = 0; probes <= mask; probes++) {
    Slot s = load_slot(&table[idx]);
    if (s.key == kEmpty) {
      if (*(volatile uint32_t*)&ctr->abort) return 0;
      unsigned long long old = atomicCAS(&table[idx].key, kEmpty, mine);
      if (old == kEmpty) return claim_new(ctr, newlist, idx, new_limit) ? provisional(table, idx, (uint32_t)r, 0xFFFFFFFFu) : 0;
      s.key = old;
      s.minrow = 0xFFFFFFFFu;
      s.nbin = 0xFFFFFFFFu;
    }
    if ((uint32_t)(s.key >> 32) == tag && rows_equal(kc, r, (int64_t)(uint32_t)s.key))
      return s.bin() ? (int32_t)s.bin() : provisional(table, idx, (uint32_t)r, s.minrow);
    idx = (idx + 1) & mask;
  }
  ctr->abort = 1;
  return 0;
}

// A warp owns 128-row tiles, lane l takes rows l, 32+l, 64+l, 96+l (every column load of the warp is contiguous).
// The four hashes, then the four home-slot probes, then the four verifications are each issued together.
__global__ void __launch_bounds__(256) gb_round_byref(KeyCols kc, const uint8_t* __restrict__ filter,
                                                      int32_t* __restrict__ ikey, Slot* table, uint32_t mask,
                                                      int64_t lo, int64_t hi, uint32_t* newlist, GbCounters* ctr,
                                                      uint32_t new_limit) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t bmask = mask >> 1;
  for (int64_t tb = lo + gw * 128; tb < hi; tb += nwarps * 128) {
    if (*(volatile uint32_t*)&ctr->abort) break;  // the round is out of room and will be redone
    int64_t r[4];
    bool live[4];
    unsigned long long h[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      r[i] = tb + i * 32 + lane;
      live[i] = r[i] < hi && (!filter || filter[r[i]]);
      h[i] = live[i] ? hash_row(kc, r[i]) : 0ull;
    }
    uint32_t idx[4];
    Slot s[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      idx[i] = ((uint32_t)h[i] & bmask) * 2;
      if (live[i]) s[i] = load_slot(&table[idx[i]]);
    }
    unsigned long long diff[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const bool cand = live[i] && s[i].key != kEmpty && (uint32_t)(s[i].key >> 32) == (uint32_t)(h[i] >> 32);
      diff[i] = cand ? rows_diff(kc, r[i], (int64_t)(uint32_t)s[i].key) : 1ull;
    }
    int32_t o[4];
    uint32_t slow = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      o[i] = 0;
      if (live[i]) {
        if (diff[i] == 0 && s[i].bin()) o[i] = (int32_t)s[i].bin();
        else slow |= 1u << i;
      }
    }
    while (slow) {  // one call site for the general path (it is inlined: KeyCols must stay in the parameter bank)
      const int i = __ffs(slow) - 1;
      slow &= slow - 1;
      const int64_t ri = i == 0 ? r[0] : i == 1 ? r[1] : i == 2 ? r[2] : r[3];
      const unsigned long long hi_ = i == 0 ? h[0] : i == 1 ? h[1] : i == 2 ? h[2] : h[3];
      const int32_t v = resolve_byref(kc, table, mask, ri, (uint32_t)(hi_ >> 32), ((uint32_t)hi_ & bmask) * 2, newlist, ctr, new_limit);
      if (i == 0) o[0] = v; else if (i == 1) o[1] = v; else if (i == 2) o[2] = v; else o[3] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (r[i] < hi) ikey[r[i]] = o[i];
  }
}

// ---- packed-row path: multi-key / string rows up to 64 bytes ---------------------------------------------------
// The key columns are first packed into one row-major buffer of W = 16, 32, 48 or 64 bytes per row (what the
// reference does too: "multikey first packs rows").  Hashing and the bytewise verification then work on fixed-size
// vector loads — one 128-bit load per 16 bytes, the representative row of a candidate slot is W contiguous bytes —
// instead of per-column loops over a runtime descriptor (measured: 1850 instructions per row for the generic
// kernel on an (int64, S16, int32) key).
// one launch per key column: packed[r*W + off .. + sz) = column[r]; G = bytes moved per load/store (a power of two
// dividing the item size, the column's base alignment and the destination offset)
template <int G>
__global__ void __launch_bounds__(256) gb_pack_column(const uint8_t* __restrict__ col, int sz, int64_t n, int W, int off,
                                                      uint8_t* __restrict__ packed) {
  typedef typename std::conditional<G == 16, uint4, typename std::conditional<G == 8, unsigned long long,
          typename std::conditional<G == 4, uint32_t, typename std::conditional<G == 2, uint16_t, uint8_t>::type>::type>::type>::type T;
  const int per = sz / G;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const T* src = (const T*)(col + r * (int64_t)sz);
    T* dst = (T*)(packed + r * (int64_t)W + off);
    for (int j = 0; j < per; j++) dst[j] = src[j];
  }
}

// Key rows of at most 8 bytes in total -- two or more small columns ((int32, int32), (int16, int32, int8) ...) or one
// column of 3, 5, 6 or 7 bytes -- are concatenated into ONE zero-padded 64-bit key per row, and the grouping then runs on
// the single-key kernels (200 G rows/s against 24 G rows/s for the packed-row route at 1 M distinct pairs).
__global__ void __launch_bounds__(256) gb_compose_u64(KeyCols kc, int64_t n, unsigned long long* __restrict__ out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long v = 0;
    int off = 0;
    for (int c = 0; c < kc.ncols; c++) {
      const int sz = kc.itemsize[c];
      const uint8_t* p = kc.ptr[c] + r * (int64_t)sz;
      unsigned long long x = 0;
      switch (sz) {
        case 8: x = *(const unsigned long long*)p; break;
        case 4: x = *(const uint32_t*)p; break;
        case 2: x = *(const uint16_t*)p; break;
        case 1: x = *p; break;
        default:
          for (int k = 0; k < sz; k++) x |= (unsigned long long)p[k] << (8 * k);
      }
      v |= x << (8 * off);
      off += sz;
    }
    out[r] = v;
  }
}

template <int W16>
struct PackedRow {
  uint4 w[W16];
};
// rowsz = bytes a row occupies in memory.  16 * W16: the packed buffer (or a column whose items are 16-byte multiples),
// 128-bit loads.  Anything smaller: ONE key column read in place (S10, S12, U3 ...), rows at arbitrary byte offsets of a
// 4-byte aligned base; the row is assembled from aligned 32-bit words with funnel shifts and zero-padded, which is
// exactly what the packing pass would have written -- without its extra read and write of the whole column.
template <int W16>
__device__ __forceinline__ PackedRow<W16> load_packed(const uint8_t* packed, int64_t r, int rowsz) {
  PackedRow<W16> x;
  if (rowsz == 16 * W16) {
    const uint4* p = (const uint4*)(packed + r * (int64_t)(16 * W16));
#pragma unroll
    for (int j = 0; j < W16; j++) x.w[j] = p[j];
    return x;
  }
  const int64_t b = r * (int64_t)rowsz;
  const uint32_t* wp = (const uint32_t*)(packed + (b & ~3ll));
  const int lead = (int)(b & 3);  // bytes of the first word that belong to the previous row
  uint32_t o[4 * W16];
#pragma unroll
  for (int j = 0; j < 4 * W16; j++) {
    uint32_t v = 0;
    const int have = rowsz - 4 * j;  // bytes of the row from output word j on
    if (have > 0) {
      const uint32_t lo = wp[j];
      const int need = have < 4 ? have : 4;
      const uint32_t hi = (lead + need > 4) ? wp[j + 1] : 0u;  // only touched when the row really reaches into it
      v = __funnelshift_r(lo, hi, lead * 8);
      if (have < 4) v &= (1u << (8 * have)) - 1u;
    }
    o[j] = v;
  }
#pragma unroll
  for (int j = 0; j < W16; j++) x.w[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  return x;
}
template <int W16>
__device__ __forceinline__ unsigned long long hash_packed(const PackedRow<W16>& x) {
  unsigned long long h = 0x243F6A8885A308D3ull;
#pragma unroll
  for (int j = 0; j < W16; j++) {
    h = mix64(h, ((unsigned long long)x.w[j].y << 32) | x.w[j].x);
    h = mix64(h, ((unsigned long long)x.w[j].w << 32) | x.w[j].z);
  }
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return h;
}
template <int W16>
__device__ __forceinline__ bool packed_equal(const PackedRow<W16>& a, const PackedRow<W16>& b) {
  uint32_t d = 0;
#pragma unroll
  for (int j = 0; j < W16; j++) d |= (a.w[j].x ^ b.w[j].x) | (a.w[j].y ^ b.w[j].y) | (a.w[j].z ^ b.w[j].z) | (a.w[j].w ^ b.w[j].w);
  return d == 0;
}

// up to four bytes at an arbitrary byte offset of a 4-byte aligned buffer, zero-extended
__device__ __forceinline__ uint32_t load_bytes4(const uint8_t* base, int64_t off, int nbytes) {
  const uint32_t* wp = (const uint32_t*)(base + (off & ~3ll));
  const int lead = (int)(off & 3);
  const uint32_t lo = wp[0];
  const uint32_t hi = (lead + nbytes > 4) ? wp[1] : 0u;
  uint32_t v = __funnelshift_r(lo, hi, lead * 8);
  if (nbytes < 4) v &= (1u << (8 * nbytes)) - 1u;
  return v;
}
// The packed row of row r assembled straight from the key columns (what gb_pack_column would have written): used by
// the shared-memory round, whose rows then never need the packed buffer.
// Where every 32-bit word of a packed row comes from when all items are multiples of 4 bytes (int32, int64, float
// columns): column index and word inside the item, -1 for padding.  Built on the host, so the device code indexes its
// registers statically and needs one aligned 4-byte load per word.
struct WordMap {
  int8_t col[8];
  int8_t k[8];
};

template <int W16>
__device__ __forceinline__ PackedRow<W16> load_row_cols(const KeyCols& kc, int64_t r, bool whole_words, const WordMap& wm) {
  uint32_t o[4 * W16];
#pragma unroll
  for (int w = 0; w < 4 * W16; w++) o[w] = 0u;
  int off = 0;
  if (whole_words) {
#pragma unroll
    for (int w = 0; w < 4 * W16; w++) {
      const int c = wm.col[w];
      if (c >= 0) o[w] = ((const uint32_t*)(kc.ptr[c] + r * (int64_t)kc.itemsize[c]))[wm.k[w]];
    }
    PackedRow<W16> y;
#pragma unroll
    for (int j = 0; j < W16; j++) y.w[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    return y;
  }
  for (int c = 0; c < kc.ncols; c++) {
    const int sz = kc.itemsize[c];
    const int64_t b = r * (int64_t)sz;
    for (int k = 0; k < sz; k += 4) {
      const uint32_t v = load_bytes4(kc.ptr[c], b + k, sz - k < 4 ? sz - k : 4);
      const int pos = off + k, word = pos >> 2, sh = (pos & 3) * 8;
      const uint32_t lo = v << sh, hi = sh ? (v >> (32 - sh)) : 0u;
#pragma unroll
      for (int w = 0; w < 4 * W16; w++) o[w] |= (w == word ? lo : 0u) | (w == word + 1 ? hi : 0u);
    }
    off += sz;
  }
  PackedRow<W16> x;
#pragma unroll
  for (int j = 0; j < W16; j++) x.w[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  return x;
}

// The packing pass in one launch: a thread assembles its row from the columns and stores it whole (the per-column launches
// of gb_pack_column each touch every 32-byte sector of the packed rows with a partial store, after a memset for the padding).
template <int W16>
__global__ void __launch_bounds__(256) gb_pack_rows_kernel(KeyCols kc, WordMap wm, int whole_words, int64_t a, int64_t b,
                                                           uint8_t* __restrict__ packed) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = a + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < b; r += stride) {
    const PackedRow<W16> x = load_row_cols<W16>(kc, r, whole_words != 0, wm);
    uint4* dst = (uint4*)(packed + r * (int64_t)(16 * W16));
#pragma unroll
    for (int j = 0; j < W16; j++) dst[j] = x.w[j];
  }
}

template <int W16>
__device__ __noinline__ int32_t resolve_packed(const uint8_t* packed, Slot* table, uint32_t mask, int64_t r, uint32_t tag,
                                               uint32_t idx, uint32_t* newlist, GbCounters* ctr, uint32_t new_limit, int rowsz,
                                               uint32_t* deferred_claim = nullptr) {
  const unsigned long long mine = ((unsigned long long)tag << 32) | (uint32_t)r;
  const PackedRow<W16> me = load_packed<W16>(packed, r, rowsz);
  for (uint32_t probes = 0; probes <= mask; probes++) {
    Slot s = load_slot(&table[idx]);
    if (s.key == kEmpty) {
      if (*(volatile uint32_t*)&ctr->abort) return 0;
      unsigned long long old = atomicCAS(&table[idx].key, kEmpty, mine);
      if (old == kEmpty) {
        if (deferred_claim) {  // the caller lists the slot together with its other rows' (claim_new_n)
          *deferred_claim = idx;
          return provisional(table, idx, (uint32_t)r, 0xFFFFFFFFu);
        }
        return claim_new(ctr, newlist, idx, new_limit) ? provisional(table, idx, (uint32_t)r, 0xFFFFFFFFu) : 0;
      }
      s.key = old;
      s.minrow = 0xFFFFFFFFu;
      s.nbin = 0xFFFFFFFFu;
    }
    if ((uint32_t)(s.key >> 32) == tag && packed_equal<W16>(me, load_packed<W16>(packed, (int64_t)(uint32_t)s.key, rowsz)))
      return s.bin() ? (int32_t)s.bin() : provisional(table, idx, (uint32_t)r, s.minrow);
    idx = (idx + 1) & mask;
  }
  ctr->abort = 1;
  return 0;
}

// A warp owns 32*R-row tiles; lane l takes rows l, 32+l, ... so that every load of the warp is contiguous.
template <int W16>
__global__ void __launch_bounds__(256) gb_round_packed(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ filter,
                                                       int32_t* __restrict__ ikey, Slot* table, uint32_t mask, int64_t lo,
                                                       int64_t hi, uint32_t* newlist, GbCounters* ctr, uint32_t new_limit, int rowsz) {
  constexpr int R = 2;  // rows per lane (4 measured 2.4x slower: 128 registers, two CTAs per SM)
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t bmask = mask >> 1;
  for (int64_t tb = lo + gw * (32 * R); tb < hi; tb += nwarps * (32 * R)) {
    if (*(volatile uint32_t*)&ctr->abort) break;  // the round is out of room and will be redone
    int64_t r[R];
    bool live[R];
    PackedRow<W16> me[R];
    unsigned long long h[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      r[i] = tb + i * 32 + lane;
      live[i] = r[i] < hi && (!filter || filter[r[i]]);
      if (live[i]) me[i] = load_packed<W16>(packed, r[i], rowsz);
    }
    Slot s[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      h[i] = live[i] ? hash_packed<W16>(me[i]) : 0ull;
      if (live[i]) s[i] = load_slot(&table[((uint32_t)h[i] & bmask) * 2]);
    }
    PackedRow<W16> rep[R];
    bool cand[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      cand[i] = live[i] && s[i].key != kEmpty && (uint32_t)(s[i].key >> 32) == (uint32_t)(h[i] >> 32) && s[i].bin() != 0;
      if (cand[i]) rep[i] = load_packed<W16>(packed, (int64_t)(uint32_t)s[i].key, rowsz);
    }
    uint32_t claimed[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      claimed[i] = kNoClaim;
      if (r[i] >= hi) continue;
      int32_t o = 0;
      if (live[i]) {
        if (cand[i] && packed_equal<W16>(me[i], rep[i])) o = (int32_t)s[i].bin();
        else o = resolve_packed<W16>(packed, table, mask, r[i], (uint32_t)(h[i] >> 32), ((uint32_t)h[i] & bmask) * 2, newlist, ctr, new_limit, rowsz,
                                     &claimed[i]);
      }
      ikey[r[i]] = o;
    }
    claim_new_n<R>(ctr, newlist, claimed, new_limit);
  }
}

// ---- packed rows, table in DRAM: the fast path and the general path in separate launches ---------------------
// gb_round_packed resolves a row that misses its home slot (a fifth of them at the usual load), lands on a provisional
// slot or is new inside the row loop: two or three more dependent DRAM-class accesses during which the other lanes of the
// warp wait -- a settled 231 M-row round of C3 ran at 6 G rows/s.  Here the first launch only takes the fast path -- both
// slots of the home bucket (one 32-byte load), the representative row of the slot whose tag matches, compare -- and marks
// every other row with kDeferred; the second launch compacts the marked rows of a tile and resolves them with all lanes busy.
constexpr int32_t kDeferred = INT32_MIN;  // never a provisional value while the table has at most 2^30 slots

template <int W16>
__global__ void __launch_bounds__(256) gb_round_packed_fast(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ filter,
                                                            int32_t* __restrict__ ikey, const Slot* __restrict__ table, uint32_t mask,
                                                            int64_t lo, int64_t hi, int rowsz) {
  constexpr int R = 2;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t bmask = mask >> 1;
  for (int64_t tb = lo + gw * (32 * R); tb < hi; tb += nwarps * (32 * R)) {
    int64_t r[R];
    bool live[R];
    PackedRow<W16> me[R];
    unsigned long long h[R];
    Bucket b[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      r[i] = tb + i * 32 + lane;
      live[i] = r[i] < hi && (!filter || filter[r[i]]);
      if (live[i]) me[i] = load_packed<W16>(packed, r[i], rowsz);
    }
#pragma unroll
    for (int i = 0; i < R; i++) {
      h[i] = live[i] ? hash_packed<W16>(me[i]) : 0ull;
      if (live[i]) b[i] = load_bucket(table, (uint32_t)h[i] & bmask);
    }
    PackedRow<W16> rep[R];
    uint32_t bin[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      bin[i] = 0;
      if (live[i]) {
        const uint32_t tag = (uint32_t)(h[i] >> 32);
        const uint32_t bin0 = ~(uint32_t)(b[i].w0 >> 32), bin1 = ~(uint32_t)(b[i].w1 >> 32);
        unsigned long long k = kEmpty;
        if (b[i].k0 != kEmpty && (uint32_t)(b[i].k0 >> 32) == tag && bin0) { k = b[i].k0; bin[i] = bin0; }
        else if (b[i].k1 != kEmpty && (uint32_t)(b[i].k1 >> 32) == tag && bin1) { k = b[i].k1; bin[i] = bin1; }
        if (bin[i]) rep[i] = load_packed<W16>(packed, (int64_t)(uint32_t)k, rowsz);
      }
    }
#pragma unroll
    for (int i = 0; i < R; i++) {
      if (r[i] >= hi) continue;
      int32_t o = 0;
      if (live[i]) o = (bin[i] && packed_equal<W16>(me[i], rep[i])) ? (int32_t)bin[i] : kDeferred;
      ikey[r[i]] = o;
    }
  }
}

template <int W16>
__global__ void __launch_bounds__(256) gb_round_packed_deferred(const uint8_t* __restrict__ packed, int32_t* __restrict__ ikey, Slot* table,
                                                                uint32_t mask, int64_t lo, int64_t hi, uint32_t* newlist,
                                                                GbCounters* ctr, uint32_t new_limit, int rowsz) {
  constexpr int T = 2048;  // rows per tile
  __shared__ uint32_t list[T];
  __shared__ uint32_t cnt, stop;
  const int lane = threadIdx.x & 31;
  for (int64_t tile = lo + (int64_t)blockIdx.x * T; tile < hi; tile += (int64_t)gridDim.x * T) {
    if (threadIdx.x == 0) {
      cnt = 0;
      stop = *(volatile uint32_t*)&ctr->abort;  // the round is out of room and will be redone
    }
    __syncthreads();
    if (stop) break;
#pragma unroll
    for (int k = 0; k < T / 256; k++) {
      const int64_t r = tile + k * 256 + threadIdx.x;
      const bool f = r < hi && ikey[r] == kDeferred;
      const uint32_t m = __ballot_sync(0xffffffffu, f);
      uint32_t base = 0;
      if (lane == 0 && m) base = atomicAdd(&cnt, (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (f) list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(r - tile);
    }
    __syncthreads();
    const uint32_t n = cnt;
    for (uint32_t j0 = 0; j0 < n; j0 += 256) {  // whole warps enter, so that the new-slot list is bumped once per warp
      const uint32_t j = j0 + threadIdx.x;
      uint32_t claimed[1] = {kNoClaim};
      if (j < n) {
        const int64_t r = tile + list[j];
        const unsigned long long h = hash_packed<W16>(load_packed<W16>(packed, r, rowsz));
        ikey[r] = resolve_packed<W16>(packed, table, mask, r, (uint32_t)(h >> 32), ((uint32_t)h & (mask >> 1)) * 2, newlist, ctr,
                                      new_limit, rowsz, &claimed[0]);
      }
      __syncwarp();
      claim_new_n<1>(ctr, newlist, claimed, new_limit);
    }
    __syncthreads();
  }
}

// ---- packed rows, low cardinality: every key of the table in shared memory --------------------------------
// Once the table is settled and its keys fit (U * row bytes up to about 160 KB: 10 K 16-byte keys, 5 K 32-byte keys),
// every CTA copies the U representative rows into shared memory in bin order (entry b holds the key of bin b + 1,
// so the bin needs no storage) behind a 16-bit open-addressing index, and the round is bound by the key / iKey
// streams instead of two dependent global gathers per row.  Keys missing from the snapshot (first seen in this round)
// fall back to the global table.
constexpr uint32_t kPackedSmemBytes = 224 * 1024;

__device__ __forceinline__ bool smem_index_claim(uint16_t* sidx, uint32_t slot, uint16_t val) {
  // 16-bit compare-and-swap of 0xFFFF -> val through the containing 32-bit word
  uint32_t* word = (uint32_t*)(sidx + (slot & ~1u));
  const int sh = (slot & 1u) * 16;
  uint32_t old = *(volatile uint32_t*)word;
  for (;;) {
    if (((old >> sh) & 0xFFFFu) != 0xFFFFu) return false;
    const uint32_t want = (old & ~(0xFFFFu << sh)) | ((uint32_t)val << sh);
    const uint32_t got = atomicCAS(word, old, want);
    if (got == old) return true;
    old = got;
  }
}

template <int W16>
__global__ void __launch_bounds__(1024, 1) gb_round_packed_smem(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ filter,
                                                                int32_t* __restrict__ ikey, Slot* table, uint32_t mask,
                                                                int64_t lo, int64_t hi, uint32_t* newlist, GbCounters* ctr,
                                                                uint32_t new_limit, const int32_t* __restrict__ ifirstkey,
                                                                uint32_t U, uint32_t icap, int rowsz, KeyCols kc, int from_cols, WordMap wm) {
  extern __shared__ __align__(16) unsigned char gb_smem[];
  uint4* skeys = (uint4*)gb_smem;                            // [U][W16]
  uint16_t* sidx = (uint16_t*)(skeys + (size_t)U * W16);     // [icap], 0xFFFF = empty
  for (uint32_t i = threadIdx.x; i < icap / 2; i += blockDim.x) ((uint32_t*)sidx)[i] = 0xFFFFFFFFu;
  __syncthreads();
  const uint32_t imask = icap - 1;
  for (uint32_t b = threadIdx.x; b < U; b += blockDim.x) {
    const PackedRow<W16> k = load_packed<W16>(packed, (int64_t)ifirstkey[b], rowsz);
#pragma unroll
    for (int j = 0; j < W16; j++) skeys[(size_t)b * W16 + j] = k.w[j];
    uint32_t slot = (uint32_t)hash_packed<W16>(k) & imask;
    while (!smem_index_claim(sidx, slot, (uint16_t)b)) slot = (slot + 1) & imask;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int R = W16 <= 2 ? 2 : 1;  // rows in flight per thread
  for (int64_t r0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < hi; r0 += stride * R) {
    PackedRow<W16> me[R];
    bool live[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int64_t r = r0 + i * stride;
      live[i] = r < hi && (!filter || filter[r]);
      if (live[i]) me[i] = from_cols ? load_row_cols<W16>(kc, r, from_cols == 2, wm) : load_packed<W16>(packed, r, rowsz);
    }
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int64_t r = r0 + i * stride;
      if (r >= hi) continue;
      int32_t o = 0;
      if (live[i]) {
        const unsigned long long h = hash_packed<W16>(me[i]);
        uint32_t slot = (uint32_t)h & imask;
        for (;;) {
          const uint32_t e = sidx[slot];
          if (e == 0xFFFFu) {  // not in the snapshot: the general path on the global table
            if (from_cols) {
              // this round's rows were never packed: publish this one before it can become a group's representative
              uint4* dst = (uint4*)(const_cast<uint8_t*>(packed) + r * (int64_t)(16 * W16));
#pragma unroll
              for (int j = 0; j < W16; j++) dst[j] = me[i].w[j];
              __threadfence();
            }
            o = resolve_packed<W16>(packed, table, mask, r, (uint32_t)(h >> 32), ((uint32_t)h & (mask >> 1)) * 2, newlist, ctr, new_limit, rowsz);
            break;
          }
          PackedRow<W16> k;
#pragma unroll
          for (int j = 0; j < W16; j++) k.w[j] = skeys[(size_t)e * W16 + j];
          if (packed_equal<W16>(me[i], k)) {
            o = (int32_t)(e + 1);
            break;
          }
          slot = (slot + 1) & imask;
        }
      }
      ikey[r] = o;
    }
  }
}

// ---- packed 16-byte rows, any cardinality: a global-memory snapshot with the key inline ----------------------
// The packed-row table stores a tag and the row id of a representative, so a probe is two dependent random accesses
// (slot, then the representative's row somewhere in the packed buffer).  For settled tables of 16-byte rows whose keys do
// not fit in shared memory, the finalised keys are also kept in a snapshot table of 32-byte entries {key, bin} -- one
// sector, one access per probe.  The snapshot is only an accelerator: it holds finalised keys with their final bins, it is
// extended between rounds from iFirstKey, and a row that misses it takes the exact path on the main table.
template <int W16>
struct SnapEntry {  // 32 bytes for 16-byte rows (one sector), 64 bytes for 32-byte rows
  uint4 key[W16];
  uint32_t bin;  // 0 = empty
  uint32_t pad[4 * W16 - 1];
};
static_assert(sizeof(SnapEntry<1>) == 32 && sizeof(SnapEntry<2>) == 64, "snapshot entries are whole sectors");

template <int W16>
__global__ void __launch_bounds__(256) gb_snap_build(const uint8_t* __restrict__ packed, int rowsz, const int32_t* __restrict__ ifirstkey,
                                                     uint32_t b0, uint32_t b1, SnapEntry<W16>* snap, uint32_t smask) {
  for (uint32_t b = b0 + blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += gridDim.x * blockDim.x) {
    const PackedRow<W16> k = load_packed<W16>(packed, (int64_t)ifirstkey[b], rowsz);
    uint32_t slot = (uint32_t)(hash_packed<W16>(k) >> 7) & smask;
    while (atomicCAS(&snap[slot].bin, 0u, b + 1u) != 0u) slot = (slot + 1) & smask;
#pragma unroll
    for (int j = 0; j < W16; j++) snap[slot].key[j] = k.w[j];
  }
}

template <int W16>
__global__ void __launch_bounds__(256, 3) gb_round_packed_snap(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ filter,
                                                               int32_t* __restrict__ ikey, Slot* table, uint32_t mask, int64_t lo,
                                                               int64_t hi, uint32_t* newlist, GbCounters* ctr, uint32_t new_limit,
                                                               const SnapEntry<W16>* __restrict__ snap, uint32_t smask, int rowsz,
                                                               KeyCols kc, int from_cols, WordMap wm) {
  constexpr int R = W16 == 1 ? 4 : 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < hi; r0 += stride * R) {
    PackedRow<W16> me[R];
    bool live[R];
    unsigned long long h[R];
    uint32_t slot[R];
    PackedRow<W16> ek[R];
    uint32_t ebin[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int64_t r = r0 + i * stride;
      live[i] = r < hi && (!filter || filter[r]);
      if (live[i]) me[i] = from_cols ? load_row_cols<W16>(kc, r, from_cols == 2, wm) : load_packed<W16>(packed, r, rowsz);
    }
#pragma unroll
    for (int i = 0; i < R; i++) {
      h[i] = live[i] ? hash_packed<W16>(me[i]) : 0ull;
      slot[i] = (uint32_t)(h[i] >> 7) & smask;
      if (live[i]) {
#pragma unroll
        for (int j = 0; j < W16; j++) ek[i].w[j] = snap[slot[i]].key[j];
        ebin[i] = snap[slot[i]].bin;
      }
    }
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int64_t r = r0 + i * stride;
      if (r >= hi) continue;
      int32_t o = 0;
      if (live[i]) {
        for (;;) {
          if (ebin[i] == 0u) {  // not in the snapshot: the exact path on the main table
            if (from_cols) {
              uint4* dst = (uint4*)(const_cast<uint8_t*>(packed) + r * (int64_t)(16 * W16));
#pragma unroll
              for (int j = 0; j < W16; j++) dst[j] = me[i].w[j];
              __threadfence();
            }
            o = resolve_packed<W16>(packed, table, mask, r, (uint32_t)(h[i] >> 32), ((uint32_t)h[i] & (mask >> 1)) * 2, newlist, ctr, new_limit, rowsz);
            break;
          }
          if (packed_equal<W16>(ek[i], me[i])) {
            o = (int32_t)ebin[i];
            break;
          }
          slot[i] = (slot[i] + 1) & smask;
#pragma unroll
          for (int j = 0; j < W16; j++) ek[i].w[j] = snap[slot[i]].key[j];
          ebin[i] = snap[slot[i]].bin;
        }
      }
      ikey[r] = o;
    }
  }
}

template <int W16>
__global__ void __launch_bounds__(256) gb_reinsert_packed(const uint8_t* __restrict__ packed, Slot* table, uint32_t mask,
                                                          const Slot* staged, uint32_t n, int rowsz) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    Slot s = staged[j];
    unsigned long long h = hash_packed<W16>(load_packed<W16>(packed, (int64_t)(uint32_t)s.key, rowsz));
    uint32_t idx = ((uint32_t)h & (mask >> 1)) * 2;
    while (atomicCAS(&table[idx].key, kEmpty, s.key) != kEmpty) idx = (idx + 1) & mask;
    table[idx].minrow = s.minrow;
    table[idx].nbin = s.nbin;
  }
}

// ---- per-round finalisation --------------------------------------------------------------------
// First rows of the round's new keys -> bitmap.  The row (relative to the round) is also kept next to the key's list entry,
// so that gb_assign_bins does not fetch the slot a second time (with a table in DRAM every such fetch is a DRAM-class access).
__global__ void __launch_bounds__(256) gb_mark_firstrows(const Slot* table, const uint32_t* newlist,
                                                         const GbCounters* ctr, int64_t lo, uint32_t* bitmap, uint32_t* minrows) {
  uint32_t n = ctr->newcount;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    uint32_t m = table[newlist[j]].minrow - (uint32_t)lo;
    minrows[j] = m;
    atomicOr(&bitmap[m >> 5], 1u << (m & 31));
  }
}

__global__ void __launch_bounds__(256) gb_assign_bins(Slot* table, const uint32_t* newlist, const GbCounters* ctr,
                                                      const uint32_t* bitmap, const uint32_t* prefix, uint32_t ubase,
                                                      const uint32_t* minrows) {
  uint32_t n = ctr->newcount;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    uint32_t m = minrows[j];
    uint32_t rank = prefix[m >> 5] + __popc(bitmap[m >> 5] & ((1u << (m & 31)) - 1u));
    table[newlist[j]].nbin = ~(ubase + rank + 1);
  }
}

// iFirstKey of the round's new bins, read off the bitmap in row order: the k-th set bit is bin ubase + k + 1.  Streaming
// (a thread per bitmap word writes its word's run of entries), where gb_assign_bins used to scatter one entry per new key.
__global__ void __launch_bounds__(256) gb_firstkeys_from_bitmap(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
                                                                int64_t words, int64_t lo, uint32_t ubase, int32_t* __restrict__ ifirstkey,
                                                                int64_t ifirst_sub) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
    uint32_t bits = bitmap[w];
    if (!bits) continue;
    uint32_t at = ubase + prefix[w];
    const int64_t row0 = lo + (w << 5) - ifirst_sub;
    while (bits) {
      ifirstkey[at++] = (int32_t)(row0 + (__ffs(bits) - 1));
      bits &= bits - 1;
    }
  }
}

// `fz.values` != nullptr (fused sum): the provisional rows of a fused round were not summed by the round kernel
// (their bin was unknown); they are summed here, as their bin is looked up.
__global__ void __launch_bounds__(256) gb_fixup(int32_t* __restrict__ ikey, const Slot* __restrict__ table, int64_t lo,
                                                int64_t hi, int32_t bin_sub, GbFuse fz) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t r = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; r < hi; r += stride) {
    if (r + 3 < hi) {
      int4 v = *(const int4*)(ikey + r);
      int32_t o[4] = {v.x, v.y, v.z, v.w};
      bool any = false;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (o[i] < 0) {
          o[i] = (int32_t)table[-(o[i] + 1)].bin() - bin_sub;
          if (fz.values) fuse_add(fz, o[i], load_raw1(fz.values, fz.vsz, r + i));
          any = true;
        } else if (bin_sub && o[i] > 0) {
          o[i] -= bin_sub;
          any = true;
        }
      }
      if (any) *(int4*)(ikey + r) = make_int4(o[0], o[1], o[2], o[3]);
    } else {
      for (int64_t q = r; q < hi; q++) {
        int32_t v = ikey[q];
        if (v < 0) {
          v = (int32_t)table[-(v + 1)].bin() - bin_sub;
          ikey[q] = v;
          if (fz.values) fuse_add(fz, v, load_raw1(fz.values, fz.vsz, q));
        } else if (bin_sub && v > 0) {
          ikey[q] = v - bin_sub;
        }
      }
    }
  }
}

// ---- resize: stage finalised entries, clear, reinsert --------------------------------------------
__global__ void __launch_bounds__(256) gb_stage_entries(const Slot* table, uint32_t nslots, Slot* staged, GbCounters* ctr) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += gridDim.x * blockDim.x) {
    Slot s = table[i];
    if (s.bin()) {
      if (i == nslots - 1) s.key = kEmpty;  // the extra slot holds the key that equals the empty marker
      staged[aggregated_inc(&ctr->total)] = s;
    }
  }
}

__global__ void __launch_bounds__(256) gb_reinsert_direct(Slot* table, uint32_t mask, const Slot* staged, uint32_t n) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    Slot s = staged[j];
    if (s.key == kEmpty) {
      s.key = 0;
      table[mask + 1] = s;
      continue;
    }
    uint32_t idx = (hash_u64(s.key) & (mask >> 1)) * 2;
    while (atomicCAS(&table[idx].key, kEmpty, s.key) != kEmpty) idx = (idx + 1) & mask;
    table[idx].minrow = s.minrow;
    table[idx].nbin = s.nbin;
  }
}

__global__ void __launch_bounds__(256) gb_reinsert_byref(KeyCols kc, Slot* table, uint32_t mask, const Slot* staged,
                                                         uint32_t n) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    Slot s = staged[j];
    unsigned long long h = hash_row(kc, (int64_t)(uint32_t)s.key);
    uint32_t idx = ((uint32_t)h & (mask >> 1)) * 2;
    while (atomicCAS(&table[idx].key, kEmpty, s.key) != kEmpty) idx = (idx + 1) & mask;
    table[idx].minrow = s.minrow;
    table[idx].nbin = s.nbin;
  }
}

__global__ void __launch_bounds__(256) gb_fill_partition_ids(int32_t* pid, int64_t n, const int64_t* cutoffs, int ncut) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    int lo = 0, hi = ncut;  // first cutoff > r
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cutoffs[mid] > r) hi = mid; else lo = mid + 1;
    }
    pid[r] = lo;
  }
}

__global__ void __launch_bounds__(256) gb_count_partition_uniques(const int32_t* ifirstkey, int64_t U, const int32_t* pid,
                                                                  uint32_t* ucount) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < U; b += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&ucount[pid[ifirstkey[b]]], 1u);
}
__global__ void __launch_bounds__(256) gb_relative_firstkeys(int32_t* ifirstkey, int64_t U, const int32_t* pid,
                                                             const int64_t* cutoffs) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < U; b += (int64_t)gridDim.x * blockDim.x) {
    int32_t row = ifirstkey[b];
    int32_t p = pid[row];
    ifirstkey[b] = row - (int32_t)(p ? cutoffs[p - 1] : 0);
  }
}
__global__ void __launch_bounds__(256) gb_renumber_partitions(int32_t* ikey, int64_t n, const int32_t* pid,
                                                              const uint32_t* ubase) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    int32_t v = ikey[r];
    if (v > 0) ikey[r] = v - (int32_t)ubase[pid[r]];
  }
}


// ---- small inputs: the whole grouping in ONE cooperative launch -----------------------------------------------------
// Below kSmallRows rows the round schedule's host round trips (a counter read-back and two stream synchronisations per
// round, 140-170 us per call) cost more than the work.  One cooperative kernel does it all, phases separated by grid
// barriers: clear the table and the first-row bitmap | insert every row (CAS claim, atomicMin of the row; the row keeps
// -(slot + 1)) | mark every claimed slot's first row in the bitmap | popcount-scan the bitmap (per-CTA totals, then the
// prefix of every word) | rank each slot by its first row = its bin, write iFirstKey | rewrite the rows.  The table it
// leaves behind has the standard format, so rtb_ismember32 probes it like any other.
constexpr int64_t kSmallRows = 1 << 20;
constexpr int kSmallPartials = 4096;  // upper bound of the cooperative grid (148 SMs x at most 4 CTAs)

struct GbSmallArgs {
  const void* keys;
  const uint8_t* filter;
  int32_t* ikey;
  int32_t* ifirstkey;
  Slot* table;
  uint32_t* bitmap;   // one bit per row
  uint32_t* prefix;   // set bits before each bitmap word
  uint32_t* partial;  // per-CTA popcounts
  uint32_t* result;   // [0] unique count, [1] overflow flag
  int64_t n, max_unique;
  uint32_t mask;      // capacity - 1
};

__device__ __forceinline__ uint32_t block_sum_256(uint32_t v, uint32_t* sh /* 9 */) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < 8; i++) t += sh[i];
    sh[8] = t;
  }
  __syncthreads();
  const uint32_t r = sh[8];
  __syncthreads();
  return r;
}

template <int ISZ>
__global__ void __launch_bounds__(256) gb_small_kernel(GbSmallArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t sh[16];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  const uint32_t cap = a.mask + 1, bmask = a.mask >> 1;
  const int64_t words = (a.n + 31) >> 5;
  // 0. clear
  for (int64_t i = tid; i <= cap; i += nthreads) {
    Slot e;
    e.key = kEmpty; e.minrow = 0xFFFFFFFFu; e.nbin = 0xFFFFFFFFu;
    a.table[i] = e;
  }
  for (int64_t w = tid; w < words; w += nthreads) a.bitmap[w] = 0;
  grid.sync();
  // 1. insert
  for (int64_t r = tid; r < a.n; r += nthreads) {
    int32_t o = 0;
    if (!a.filter || a.filter[r]) {
      const unsigned long long k = load_key1<ISZ>(a.keys, r);
      uint32_t idx;
      if (k == kEmpty) {  // the key that equals the empty marker lives in the extra slot
        idx = cap;
        atomicCAS(&a.table[idx].key, kEmpty, 0ull);
      } else {
        idx = (hash_u64(k) & bmask) * 2;
        for (;;) {
          const unsigned long long cur = *(volatile unsigned long long*)&a.table[idx].key;
          if (cur == k) break;
          if (cur == kEmpty) {
            const unsigned long long old = atomicCAS(&a.table[idx].key, kEmpty, k);
            if (old == kEmpty || old == k) break;
          }
          idx = (idx + 1) & a.mask;
        }
      }
      atomicMin(&a.table[idx].minrow, (uint32_t)r);
      o = -(int32_t)(idx + 1);
    }
    a.ikey[r] = o;
  }
  grid.sync();
  // 2. first rows -> bitmap
  for (int64_t i = tid; i <= cap; i += nthreads) {
    const Slot e = a.table[i];
    if (e.key != kEmpty) atomicOr(&a.bitmap[e.minrow >> 5], 1u << (e.minrow & 31));
  }
  grid.sync();
  // 3. popcount scan: every CTA owns a contiguous range of words
  const int64_t per = (words + gridDim.x - 1) / gridDim.x;
  const int64_t w0 = (int64_t)blockIdx.x * per, w1 = w0 + per < words ? w0 + per : words;
  {
    uint32_t c = 0;
    for (int64_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) c += __popc(a.bitmap[w]);
    c = block_sum_256(c, sh);
    if (threadIdx.x == 0) a.partial[blockIdx.x] = c;
  }
  grid.sync();
  {
    uint32_t before = 0, total = 0;
    for (uint32_t j = threadIdx.x; j < gridDim.x; j += blockDim.x) {
      const uint32_t p = a.partial[j];
      total += p;
      if (j < blockIdx.x) before += p;
    }
    before = block_sum_256(before, sh);
    total = block_sum_256(total, sh);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      a.result[0] = total;
      a.result[1] = total > (uint32_t)a.max_unique ? 1u : 0u;
    }
    uint32_t run = before;
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int64_t b0 = w0; b0 < w1; b0 += blockDim.x) {
      const int64_t w = b0 + threadIdx.x;
      const uint32_t v = w < w1 ? (uint32_t)__popc(a.bitmap[w]) : 0u;
      uint32_t inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (lane == 31) sh[wp] = inc;
      __syncthreads();
      uint32_t base = 0, chunk = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (i < wp) base += sh[i];
        chunk += sh[i];
      }
      if (w < w1) a.prefix[w] = run + base + inc - v;
      run += chunk;
      __syncthreads();
    }
  }
  grid.sync();
  // 4. bins
  const bool overflow = a.result[1] != 0;
  if (!overflow) {
    for (int64_t i = tid; i <= cap; i += nthreads) {
      const Slot e = a.table[i];
      if (e.key != kEmpty) {
        const uint32_t m = e.minrow;
        const uint32_t bin = a.prefix[m >> 5] + __popc(a.bitmap[m >> 5] & ((1u << (m & 31)) - 1u)) + 1u;
        a.table[i].nbin = ~bin;
        a.ifirstkey[bin - 1] = (int32_t)m;
      }
    }
  }
  grid.sync();
  // 5. rows
  if (!overflow) {
    for (int64_t r = tid; r < a.n; r += nthreads) {
      const int32_t v = a.ikey[r];
      if (v < 0) a.ikey[r] = (int32_t)a.table[-(v + 1)].bin();
    }
  }
}


// ---- adopting a seed (row shards, riptable_b200/dist.py) --------------------------------------------------------------
// A shard that has grouped a prefix of its rows under LOCAL bins receives the seed: u0 distinct keys whose bins are
// 1..u0 in seed order (the bins shard 0 gave them = the global bins).  The table is re-labelled in place: seed keys found
// in it take their seed bin, seed keys it lacks are inserted (first row unknown: 0xFFFFFFFF), local keys the seed lacks
// ("late" keys) are numbered u0 + 1 .. in their local first-appearance order.  `map` (old bin -> new bin) lets the caller
// re-index the prefix's iKey; iFirstKey and the accumulators are permuted here.
template <int ISZ>
__global__ void __launch_bounds__(256) gb_adopt_probe_kernel(const void* __restrict__ seed, int64_t u0, Slot* table, uint32_t mask,
                                                             int32_t* __restrict__ map) {
  const uint32_t bmask = mask >> 1;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < u0; j += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = load_key1<ISZ>(seed, j);
    const uint32_t gbin = (uint32_t)(j + 1);
    uint32_t idx;
    if (k == kEmpty) {
      idx = mask + 1;
      const unsigned long long old = atomicCAS(&table[idx].key, kEmpty, 0ull);
      if (old == kEmpty) {
        table[idx].nbin = ~gbin;  // fresh: minrow stays 0xFFFFFFFF
        continue;
      }
    } else {
      idx = (hash_u64(k) & bmask) * 2;
      bool fresh = false;
      for (;;) {
        const unsigned long long cur = *(volatile unsigned long long*)&table[idx].key;
        if (cur == k) break;
        if (cur == kEmpty) {
          const unsigned long long old = atomicCAS(&table[idx].key, kEmpty, k);
          if (old == kEmpty) { fresh = true; break; }
          if (old == k) break;
        }
        idx = (idx + 1) & mask;
      }
      if (fresh) {
        table[idx].nbin = ~gbin;
        continue;
      }
    }
    map[table[idx].bin()] = (int32_t)gbin;  // a local key: its slot is re-labelled by gb_adopt_relabel_kernel
  }
}
__global__ void __launch_bounds__(256) gb_adopt_late_flags_kernel(const int32_t* __restrict__ map, int64_t u_loc, uint32_t* __restrict__ flags) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < u_loc; b += (int64_t)gridDim.x * blockDim.x)
    flags[b] = map[b + 1] == 0 ? 1u : 0u;
}
__global__ void __launch_bounds__(256) gb_adopt_late_bins_kernel(int32_t* __restrict__ map, int64_t u_loc, const uint32_t* __restrict__ flags,
                                                                 const uint32_t* __restrict__ excl, uint32_t u0) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < u_loc; b += (int64_t)gridDim.x * blockDim.x)
    if (flags[b]) map[b + 1] = (int32_t)(u0 + 1 + excl[b]);
}
__global__ void __launch_bounds__(256) gb_adopt_relabel_kernel(Slot* table, uint32_t nslots, const int32_t* __restrict__ map) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += gridDim.x * blockDim.x) {
    const Slot e = table[i];
    if (e.key != kEmpty && e.minrow != 0xFFFFFFFFu) table[i].nbin = ~(uint32_t)map[e.bin()];
  }
}
// out[map[b] - shift] = in[b] for b in [first, u_loc]; item = 4 or 8 bytes
template <typename T>
__global__ void __launch_bounds__(256) gb_adopt_permute_kernel(const T* __restrict__ in, T* __restrict__ out, const int32_t* __restrict__ map,
                                                               int64_t first, int64_t u_loc, int shift) {
  for (int64_t b = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= u_loc; b += (int64_t)gridDim.x * blockDim.x)
    out[map[b] - shift] = in[b - shift];
}

// ---- host side -------------------------------------------------------------------------------------
static uint32_t next_pow2_u64(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return (uint32_t)(p > (1ull << 30) ? (1ull << 30) : p);
}

// Tunables of the round schedule and the kernel choice; every one has an RTB_GB_* environment override that exists for
// measurement (profiles/), not for production use.
struct GbPlan {
  int l2_window = 0;                  // L2 access-policy window on the table (measured slower)
  int packed_rows = 1;                // packed-row kernels for multi-key / string rows (0: by-reference kernel)
  int smem_table = 1;                 // shared-memory tables for settled low-cardinality rounds
  int prefetch = 1;                   // gb_round_direct: next tile's keys in flight during the probes
  int small_path = 1;                 // inputs up to kSmallRows rows: one cooperative launch instead of the round schedule
  int fuse_prefetch = 0;              // the same for the fused (group-and-sum) form: 1 = 3 CTAs per SM with, 0 = 4 without
  double discovery_frac = 0.002;      // a round whose predecessor found more new keys than this share of its rows is a discovery round
  int ctas_per_sm = 4;                // gb_round_direct grid = 148 x this
  int64_t round0 = 1 << 20;           // rows of the first round
  double load_target = 0.25;          // table sized so projected uniques / capacity <= this
  double load_abort = 0.75;           // a round aborts when uniques / capacity would exceed this
  int ratio = 4;                      // growth of the row ranges from round to round
  double big_table_bytes = 8.0 * 1024 * 1024 * 1024;  // above this table size every round uses the slot-wise kernel
  int64_t finish_rows = 256ll << 20;  // a settled table finishes a remainder of at most this many rows in one round
  int snapshot = 1;                   // 16- / 32-byte packed rows: inline-key snapshot table for the settled rounds
  uint32_t window_slots = 1u << 22;   // the 64 MB table window (see the resize rule)
  double window_load = 0.55;          // ... which a table keeps up to this load
  double dense_load = 0.5;            // target load of a single-key table that has to leave the window
  int window_packed = 0;              // the same two rules for packed-row (multi-key) tables
  int defer_packed = 1;               // packed rows over a table beyond defer_table_bytes: fast path and general path in separate launches
  double defer_table_bytes = 128.0 * 1024 * 1024;
  double snap_fill = 2.0;             // snapshot entries per key (power of two above it)
};

static GbPlan get_plan() {
  GbPlan p;
  if (const char* e = getenv("RTB_GB_SNAPSHOT")) p.snapshot = atoi(e);
  if (const char* e = getenv("RTB_GB_SNAPFILL")) p.snap_fill = atof(e);
  if (const char* e = getenv("RTB_GB_WINDOW_LOAD")) p.window_load = atof(e);
  if (const char* e = getenv("RTB_GB_DENSE_LOAD")) p.dense_load = atof(e);
  if (const char* e = getenv("RTB_GB_WINDOW_PACKED")) p.window_packed = atoi(e);
  if (const char* e = getenv("RTB_GB_DEFER")) p.defer_packed = atoi(e);
  if (const char* e = getenv("RTB_GB_DEFER_BYTES")) p.defer_table_bytes = atof(e);
  if (const char* e = getenv("RTB_GB_FINISH")) p.finish_rows = atoll(e);
  if (const char* e = getenv("RTB_GB_BIGTABLE")) p.big_table_bytes = atof(e);
  if (const char* e = getenv("RTB_GB_L2WIN")) p.l2_window = atoi(e);
  if (const char* e = getenv("RTB_GB_DISCOVERY")) p.discovery_frac = atof(e);
  if (const char* e = getenv("RTB_GB_PACKED")) p.packed_rows = atoi(e);
  if (const char* e = getenv("RTB_GB_SMEM")) p.smem_table = atoi(e);
  if (const char* e = getenv("RTB_GB_PREFETCH")) p.prefetch = atoi(e);
  if (const char* e = getenv("RTB_GB_FUSEPF")) p.fuse_prefetch = atoi(e);
  if (const char* e = getenv("RTB_GB_SMALL")) p.small_path = atoi(e);
  if (const char* e = getenv("RTB_GB_CTAS")) p.ctas_per_sm = atoi(e);
  if (const char* e = getenv("RTB_GB_ROUND0")) p.round0 = atoll(e);
  if (const char* e = getenv("RTB_GB_LOAD")) p.load_target = atof(e);
  if (const char* e = getenv("RTB_GB_RATIO")) p.ratio = atoi(e);
  if (p.round0 < 1024) p.round0 = 1024;
  p.round0 &= ~1023ll;
  if (p.ratio < 2) p.ratio = 2;
  return p;
}

static uint32_t cap_max_for(int64_t max_unique) { return next_pow2_u64((uint64_t)(2.0 * (double)max_unique) + 1024); }

struct GbLayout {
  Slot* table;
  Slot* staged;       // max_unique slots; aliased by newlist during rounds
  uint32_t* bitmap;   // words for the largest round
  int64_t bitmap_words;
  uint32_t* prefix;
  uint32_t* scan_scratch;
  GbCounters* ctr;
  uint32_t* small_partial;  // per-CTA popcounts of the single-launch small path (kSmallPartials entries)
  int32_t* pid;       // partition ids (cutoffs only)
  int64_t* cutoffs_dev;
  uint32_t* ucount;   // per-partition unique counts, then exclusive bases
  uint8_t* packed;    // row-major packed key rows (multi-key / string rows up to 64 bytes), or nullptr
  uint8_t* snap;      // inline-key snapshot for 16- and 32-byte rows (32 / 64 bytes per entry), or nullptr
  size_t bytes_without_packed;
  size_t bytes;
};

constexpr uint32_t kSnapCapMax = 1u << 25;  // 32 Mi entries = 1 GB
static uint32_t snap_cap_for(int64_t max_unique) {
  const uint64_t want = 4ull * (uint64_t)max_unique + 1024;
  if (want >= kSnapCapMax) return kSnapCapMax;
  uint32_t c = 1024;
  while (c < want) c <<= 1;
  return c;
}

// Rows of the largest round a call over n rows can run.  The schedule depends on what the rounds find (geometric
// growth, one round that finishes a settled table's remainder, a resumed chunk that starts settled), so the first-row
// bitmap and its prefix counts are simply sized for the whole call: n / 4 bytes, touched only as far as a round reaches.
static int64_t largest_round(int64_t n, const GbPlan& p) {
  (void)p;
  return n;
}

static GbLayout layout(void* ws, size_t wsbytes, int64_t nrows, int64_t max_unique, bool cutoffs, int64_t ncut,
                       int packed_row_bytes = 0) {
  GbPlan p = get_plan();
  Arena a(ws, wsbytes);
  GbLayout L;
  uint32_t capmax = cap_max_for(max_unique);
  L.table = a.take<Slot>((size_t)capmax + 1);
  L.staged = a.take<Slot>((size_t)max_unique + 1);
  int64_t words = (largest_round(nrows, p) + 31) / 32 + 1;
  L.bitmap_words = words;
  L.bitmap = a.take<uint32_t>((size_t)words);
  L.prefix = a.take<uint32_t>((size_t)words);
  L.scan_scratch = a.take<uint32_t>(scan_u32_workspace_elems(words));
  L.ctr = a.take<GbCounters>(1);
  L.small_partial = a.take<uint32_t>(kSmallPartials);
  L.pid = cutoffs ? a.take<int32_t>((size_t)nrows) : nullptr;
  L.cutoffs_dev = cutoffs ? a.take<int64_t>((size_t)ncut) : nullptr;
  L.ucount = cutoffs ? a.take<uint32_t>((size_t)ncut) : nullptr;
  L.bytes_without_packed = a.off;
  L.packed = packed_row_bytes > 0 ? a.take<uint8_t>((size_t)nrows * packed_row_bytes + 64) : nullptr;
  L.snap = (packed_row_bytes == 16 || packed_row_bytes == 32) ? a.take<uint8_t>((size_t)snap_cap_for(max_unique) * 2 * packed_row_bytes) : nullptr;
  L.bytes = a.off;
  return L;
}

// ---- 8f.1: membership lookup against a finished table -----------------------------------------------------
// rc.IsMember32 on one numeric key column: the table is built by grouping `b` (its slots keep the FIRST row of every
// key, which is the answer ismember wants), then every row of `a` is looked up without inserting anything.
__global__ void __launch_bounds__(256) ismember_valid_kernel(const void* __restrict__ b, int dt, int64_t n, uint8_t* __restrict__ valid) {
  const int sz = dt <= RTB_U8 ? 1 : dt <= RTB_U16 ? 2 : (dt <= RTB_U32 || dt == RTB_F32) ? 4 : 8;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    bool ok = true;
    if (dt == RTB_F64) ok = (((const unsigned long long*)b)[r] & 0x7FFFFFFFFFFFFFFFull) <= 0x7FF0000000000000ull;
    else if (dt == RTB_F32) ok = (((const uint32_t*)b)[r] & 0x7FFFFFFFu) <= 0x7F800000u;
    (void)sz;
    valid[r] = ok;
  }
}

__device__ __forceinline__ void store_index(void* out, int idt, int64_t r, long long v) {
  switch (idt) {
    case RTB_I8: ((int8_t*)out)[r] = (int8_t)v; break;
    case RTB_I16: ((int16_t*)out)[r] = (int16_t)v; break;
    case RTB_I32: ((int32_t*)out)[r] = (int32_t)v; break;
    default: ((long long*)out)[r] = v;
  }
}
// two / four consecutive rows in one store
__device__ __forceinline__ void store_index2(void* out, int idt, int64_t r, long long v0, long long v1) {
  switch (idt) {
    case RTB_I8: *(uchar2*)((int8_t*)out + r) = make_uchar2((uint8_t)v0, (uint8_t)v1); break;
    case RTB_I16: *(ushort2*)((int16_t*)out + r) = make_ushort2((uint16_t)v0, (uint16_t)v1); break;
    case RTB_I32: *(int2*)((int32_t*)out + r) = make_int2((int32_t)v0, (int32_t)v1); break;
    default: *(longlong2*)((long long*)out + r) = make_longlong2(v0, v1);
  }
}

// Looks one key up from bucket `b` on (already fetched as `p`): position of its first occurrence in b, or -1.
__device__ __forceinline__ long long ismember_probe(const Slot* table, uint32_t mask, unsigned long long k, uint32_t b, Bucket p) {
  const uint32_t bmask = mask >> 1;
  if (k == kEmpty) {
    const Slot s = table[mask + 1];  // the key that equals the empty marker lives in the extra slot
    return (s.key != kEmpty && s.bin()) ? (long long)s.minrow : -1;
  }
  for (uint32_t hops = 0; hops <= bmask; hops++) {
    if (p.k0 == k) return (uint32_t)p.w0;
    if (p.k1 == k) return (uint32_t)p.w1;
    if (p.k0 == kEmpty || p.k1 == kEmpty) return -1;
    b = (b + 1) & bmask;
    p = load_bucket(table, b);
  }
  return -1;
}

// Same tiling as gb_round_direct: a warp owns 128-row tiles, every lane four keys, the next tile's keys in flight.
template <int ISZ>
__global__ void __launch_bounds__(256) ismember_lookup_kernel(const void* __restrict__ a, int64_t n, const Slot* __restrict__ table,
                                                              uint32_t mask, int base_index, int idt, long long absent,
                                                              uint8_t* __restrict__ found, void* __restrict__ index_out) {
  const int lane = threadIdx.x & 31;
  const uint32_t bmask = mask >> 1;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t step = nwarps * 128;
  int64_t tb = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 128;
  unsigned long long k[4], kn[4] = {0, 0, 0, 0};
  if (tb + 128 <= n) load_tile_keys<ISZ>(a, tb, lane, k);
  for (; tb + 128 <= n; tb += step) {
    if (tb + step + 128 <= n) load_tile_keys<ISZ>(a, tb + step, lane, kn);
    uint32_t b[4];
    Bucket p[4];
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = hash_u64(k[i]) & bmask;
#pragma unroll
    for (int i = 0; i < 4; i++) p[i] = load_bucket(table, b[i]);
    long long v[4];
    uint8_t f[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const long long pos = ismember_probe(table, mask, k[i], b[i], p[i]);
      f[i] = pos >= 0;
      v[i] = pos >= 0 ? pos + base_index : absent;
    }
    if (ISZ == 8) {
      const int64_t r0 = tb + 2 * lane, r1 = tb + 64 + 2 * lane;
      *(uchar2*)(found + r0) = make_uchar2(f[0], f[1]);
      *(uchar2*)(found + r1) = make_uchar2(f[2], f[3]);
      store_index2(index_out, idt, r0, v[0], v[1]);
      store_index2(index_out, idt, r1, v[2], v[3]);
    } else {
      const int64_t r0 = tb + 4 * lane;
      *(uchar4*)(found + r0) = make_uchar4(f[0], f[1], f[2], f[3]);
      store_index2(index_out, idt, r0, v[0], v[1]);
      store_index2(index_out, idt, r0 + 2, v[2], v[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) k[i] = kn[i];
  }
  if (tb < n) {  // the warp whose next tile is the partial one
    for (int64_t r = tb + lane; r < n; r += 32) {
      const unsigned long long kk = load_key1<ISZ>(a, r);
      const uint32_t bb = hash_u64(kk) & bmask;
      const long long pos = ismember_probe(table, mask, kk, bb, load_bucket(table, bb));
      found[r] = pos >= 0;
      store_index(index_out, idt, r, pos >= 0 ? pos + base_index : absent);
    }
  }
}

}  // namespace rtb

using namespace rtb;

extern "C" size_t rtb_groupby_workspace_bytes_rows(int64_t nrows, int64_t max_unique, int64_t row_bytes, int has_cutoffs) {
  if (nrows < 0 || max_unique < 0 || row_bytes < 0) return 0;
  if (has_cutoffs) row_bytes += 4;
  int W = (int)((row_bytes + 15) / 16 * 16);
  if (W > 64) W = 0;
  GbLayout L = layout(nullptr, 0, nrows, max_unique, has_cutoffs != 0, has_cutoffs ? nrows : 0, W);
  return L.bytes + 256;
}

extern "C" size_t rtb_groupby_workspace_bytes(int64_t nrows, int64_t max_unique, int nkeys, int has_cutoffs) {
  (void)nkeys;
  if (nrows < 0 || max_unique < 0) return 0;
  GbLayout L = layout(nullptr, 0, nrows, max_unique, has_cutoffs != 0, has_cutoffs ? nrows : 0);
  return L.bytes + 256;
}

// Experiment, OFF by default (RTB_GB_L2WIN=1): pin the hash table in L2 with an access-policy window so that the
// key / iKey streams stop evicting table lines (ncu: 16 B/row of DRAM traffic against 12 algorithmic, L2 hit rate
// 56 %).  Measured on C2: steady rounds drop from 200 to 165 G rows/s (the persisting carve-out takes L2 ways away
// from the streams, and the round is L1TEX-bound, not miss-bound) -- kept only so the measurement can be repeated.
static void set_table_l2_window(cudaStream_t stream, const void* table, size_t bytes, bool on) {
  static int max_window = -1, max_persist = -1;
  if (max_window < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    if (max_persist > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
    cudaGetLastError();
  }
  if (max_window <= 0 || max_persist <= 0) return;
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  if (on) {
    size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    v.accessPolicyWindow.base_ptr = const_cast<void*>(table);
    v.accessPolicyWindow.num_bytes = win;
    v.accessPolicyWindow.hitRatio = win <= (size_t)max_persist ? 1.0f : (float)max_persist / (float)win;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  }
  cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v);
  cudaGetLastError();
}

static uint32_t packed_smem_icap(int64_t U) {
  uint32_t c = 1024;
  while ((int64_t)c < 2 * U) c <<= 1;
  return c;
}
static size_t packed_smem_bytes(int64_t U, int W) { return (size_t)U * W + (size_t)packed_smem_icap(U) * 2 + 16; }

// The round loop shared by the one-shot entry point and the resumable (chunked) one.  `st` carries the table state
// (capacity, unique count, discovery rate) from one chunk to the next; the table itself lives at the start of the
// workspace, whose layout depends only on max_unique.  `row_offset` is added to the first rows written to iFirstKey.
static int gb_core(const rtb_column* keys, int nkeys, int64_t nrows, const uint8_t* filter, const int64_t* cutoffs_host,
                   int64_t ncutoffs, int32_t* ikey_out, int32_t* ifirstkey_out, int64_t max_unique,
                   int64_t* unique_count_host, int64_t* ucutoffs_host, int64_t* needed_unique_host, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream, rtb_gb_state* st, int64_t row_offset,
                   const GbFuse* fuse = nullptr) {
  const bool has_cut = cutoffs_host != nullptr && ncutoffs > 0;
  GbPlan plan = get_plan();
  static const bool trace = getenv("RTB_GB_TRACE") != nullptr;
  // key layout: direct (one numeric column), packed rows (anything else up to 64 bytes per row), generic by-reference
  bool direct = !has_cut && nkeys == 1 && (keys[0].itemsize == 1 || keys[0].itemsize == 2 || keys[0].itemsize == 4 ||
                                           keys[0].itemsize == 8) &&
                ((uintptr_t)keys[0].data % vec4_align(keys[0].itemsize)) == 0 && (!filter || ((uintptr_t)filter % 4) == 0);
  int64_t row_bytes = has_cut ? 4 : 0;
  for (int c = 0; c < nkeys; c++) row_bytes += keys[c].itemsize;
  int W = (int)((row_bytes + 15) / 16 * 16);
  if (direct || W > 64 || !plan.packed_rows) W = 0;
  // rows of at most 8 bytes that are not one 1/2/4/8-byte column: composed into one 64-bit key (the packed-row buffer
  // holds the composed column), then the single-key kernels
  bool compose = !direct && !has_cut && W == 16 && row_bytes <= 8 && (!filter || ((uintptr_t)filter % 4) == 0);
  for (int c = 0; c < nkeys && compose; c++) {
    const int sz = keys[c].itemsize;
    if ((sz == 8 || sz == 4 || sz == 2) && ((uintptr_t)keys[c].data % sz) != 0) compose = false;
  }
  GbLayout L = layout(workspace, workspace_bytes, nrows, max_unique, has_cut, ncutoffs, W);
  if (W && L.bytes > workspace_bytes) {  // a caller that sized the workspace without the row width: generic kernel
    W = 0;
    L = layout(workspace, workspace_bytes, nrows, max_unique, has_cut, ncutoffs, 0);
  }
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "MultiKeyGroupBy32: workspace too small (%zu < %zu)",
          workspace_bytes, L.bytes);
  if (W == 0) compose = false;  // no packed-row buffer to hold the composed column
  const uint32_t capmax = cap_max_for(max_unique);
  uint32_t* newlist = (uint32_t*)L.staged;
  KeyCols kc;
  make_keycols(keys, nkeys, &kc);
  if (has_cut) {
    // cutoffs == an extra leading key column holding the partition id (tests/test_pgroupby.py:27-41)
    RTB_CUDA(cudaMemcpyAsync(L.cutoffs_dev, cutoffs_host, sizeof(int64_t) * ncutoffs, cudaMemcpyHostToDevice, stream));
    gb_fill_partition_ids<<<grid_for(nrows, 256, 1), 256, 0, stream>>>(L.pid, nrows, L.cutoffs_dev, (int)ncutoffs);
    RTB_LAUNCH_CHECK();
    kc.ptr[nkeys] = (const uint8_t*)L.pid;
    kc.itemsize[nkeys] = 4;
    kc.chunk[nkeys] = 4;
    kc.ncols = nkeys + 1;
  }
  int rowsz = W;  // bytes per row as the kernels read them
  rtb_column dkey = keys[0];  // the column the single-key kernels read
  if (compose) {
    gb_compose_u64<<<grid_for(nrows, 256, 2, 16), 256, 0, stream>>>(kc, nrows, (unsigned long long*)L.packed);
    RTB_LAUNCH_CHECK();
    dkey.data = L.packed;
    dkey.itemsize = 8;
    direct = true;
    W = 0;
  } else if (W && kc.ncols == 1 && kc.itemsize[0] == W && ((uintptr_t)kc.ptr[0] % 16) == 0) {
    // one column whose items already are the packed rows (S16, S32, U4 ...): probe it in place
    L.packed = const_cast<uint8_t*>(kc.ptr[0]);
  } else if (W && kc.ncols == 1 && ((uintptr_t)kc.ptr[0] % 4) == 0) {
    // one column of any other item size (S10, S12, U3 ...): also in place, rows assembled from unaligned words
    L.packed = const_cast<uint8_t*>(kc.ptr[0]);
    rowsz = kc.itemsize[0];
  }
  // several columns: rows are packed round by round, right before a packed-row kernel reads them -- the shared-memory
  // rounds (most rows of a low-cardinality multi-key grouping) assemble their rows from the columns and skip the pass
  const bool lazy_pack = W != 0 && L.packed != nullptr && !(kc.ncols == 1 && ((uintptr_t)kc.ptr[0] % 4) == 0);
  bool cols_aligned = true;
  for (int c = 0; c < kc.ncols; c++) cols_aligned = cols_aligned && ((uintptr_t)kc.ptr[c] % 4) == 0;
  WordMap wm;
  bool whole_words_all = W != 0 && W <= 32;
  for (int w = 0; w < 8; w++) wm.col[w] = -1, wm.k[w] = 0;
  {
    int word = 0;
    for (int c = 0; c < kc.ncols && whole_words_all; c++) {
      if (kc.itemsize[c] % 4 != 0) { whole_words_all = false; break; }
      for (int k = 0; k < kc.itemsize[c] / 4 && word < 8; k++, word++) wm.col[word] = (int8_t)c, wm.k[word] = (int8_t)k;
    }
  }
  auto pack_range = [&](int64_t a, int64_t b) -> int {
    if (b <= a) return RTB_OK;
    if (W <= 32 && cols_aligned) {  // rows of up to 32 bytes from 4-byte aligned columns: one launch, whole rows
      const int grid = grid_for(b - a, 256, 2, 16);
      if (W == 16) gb_pack_rows_kernel<1><<<grid, 256, 0, stream>>>(kc, wm, whole_words_all ? 1 : 0, a, b, L.packed);
      else gb_pack_rows_kernel<2><<<grid, 256, 0, stream>>>(kc, wm, whole_words_all ? 1 : 0, a, b, L.packed);
      RTB_LAUNCH_CHECK();
      return RTB_OK;
    }
    uint8_t* dst = L.packed + a * (int64_t)W;
    if (row_bytes < W) RTB_CUDA(cudaMemsetAsync(dst, 0, (size_t)(b - a) * W, stream));  // zero the padding
    int off = 0;
    for (int c = 0; c < kc.ncols; c++) {
      const int sz = kc.itemsize[c];
      int g = kc.chunk[c];
      while (g > 1 && ((off % g) != 0 || ((a * (int64_t)sz) % g) != 0)) g >>= 1;
      const int grid = grid_for(b - a, 256, 2, 16);
      const uint8_t* src = kc.ptr[c] + a * (int64_t)sz;
      switch (g) {
        case 16: gb_pack_column<16><<<grid, 256, 0, stream>>>(src, sz, b - a, W, off, dst); break;
        case 8: gb_pack_column<8><<<grid, 256, 0, stream>>>(src, sz, b - a, W, off, dst); break;
        case 4: gb_pack_column<4><<<grid, 256, 0, stream>>>(src, sz, b - a, W, off, dst); break;
        case 2: gb_pack_column<2><<<grid, 256, 0, stream>>>(src, sz, b - a, W, off, dst); break;
        default: gb_pack_column<1><<<grid, 256, 0, stream>>>(src, sz, b - a, W, off, dst);
      }
      RTB_LAUNCH_CHECK();
      off += sz;
    }
    return RTB_OK;
  };

  GbFuse fz_none;
  memset(&fz_none, 0, sizeof(fz_none));
  // fused sum: accumulators are zeroed as bins come into existence (bin 0 now, new bins after every round), so the
  // cost follows the bins actually found, not max_unique
  // (a resumed chunk keeps what earlier chunks accumulated; a chunk without values only creates -- and zeroes -- bins)
  if (fuse && st->rounds == 0) RTB_CUDA(cudaMemsetAsync(fuse->acc, 0, sizeof(unsigned long long), stream));
  const bool fuse_values = fuse && fuse->values != nullptr;
  const bool fuse_kernel_ok = fuse_values && (fuse->vsz == 8 || fuse->vsz == 4) && ((uintptr_t)fuse->values % 16) == 0;
  // ---- small one-shot inputs: a single cooperative launch (see gb_small_kernel)
  if (direct && plan.small_path && st->rounds == 0 && row_offset == 0 && nrows <= kSmallRows) {
    const uint32_t want_cap = next_pow2_u64((uint64_t)(2 * nrows < 1024 ? 1024 : 2 * nrows));
    static int coop = -1, occ8 = 0, occ4 = 0, occ2 = 0, occ1 = 0;
    if (coop < 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ8, gb_small_kernel<8>, 256, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, gb_small_kernel<4>, 256, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, gb_small_kernel<2>, 256, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, gb_small_kernel<1>, 256, 0);
      cudaGetLastError();
    }
    const int occ = dkey.itemsize == 8 ? occ8 : dkey.itemsize == 4 ? occ4 : dkey.itemsize == 2 ? occ2 : occ1;
    if (coop > 0 && occ > 0 && want_cap <= capmax && max_unique >= 1) {
      // one row per thread while the machine has room (the insert and rewrite phases are chains of dependent L2 round trips:
      // few threads with many rows each made a 10 K-row call take 87 us); never more CTAs than can be resident at once
      int64_t ctas = (nrows + 255) / 256;
      const int64_t resident = (int64_t)kSMs * (occ < 4 ? occ : 4);
      if (ctas > resident) ctas = resident;
      if (ctas < 1) ctas = 1;
      GbSmallArgs sa;
      sa.keys = dkey.data; sa.filter = filter; sa.ikey = ikey_out; sa.ifirstkey = ifirstkey_out; sa.table = L.table;
      sa.bitmap = L.bitmap; sa.prefix = L.prefix; sa.partial = L.small_partial; sa.result = (uint32_t*)L.ctr;
      sa.n = nrows; sa.max_unique = max_unique; sa.mask = want_cap - 1;
      void* kargs[] = {&sa};
      const void* fn = dkey.itemsize == 8 ? (const void*)gb_small_kernel<8> : dkey.itemsize == 4 ? (const void*)gb_small_kernel<4>
                       : dkey.itemsize == 2 ? (const void*)gb_small_kernel<2> : (const void*)gb_small_kernel<1>;
      prof_begin(PROF_GB_ROUND, stream);
      RTB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)ctas), dim3(256), kargs, 0, stream));
      RTB_LAUNCH_CHECK();
      prof_end(PROF_GB_ROUND, stream, (uint64_t)nrows);
      // the two result words come back through a page-locked buffer (a pageable target costs the driver a staging copy)
      static thread_local uint32_t* pinned = nullptr;
      if (!pinned && cudaHostAlloc((void**)&pinned, 64, cudaHostAllocDefault) != cudaSuccess) {
        pinned = nullptr;
        cudaGetLastError();
      }
      uint32_t res_local[2] = {0, 0};
      uint32_t* res = pinned ? pinned : res_local;
      RTB_CUDA(cudaMemcpyAsync(res, L.ctr, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      RTB_CUDA(cudaStreamSynchronize(stream));
      if (res[1]) {
        if (needed_unique_host) *needed_unique_host = (int64_t)res[0];
        set_error("MultiKeyGroupBy32: more than %lld unique keys; retry with max_unique >= %u", (long long)max_unique, res[0]);
        return RTB_ERR_WORKSPACE;
      }
      const int64_t Us = (int64_t)res[0];
      if (fuse && fuse->values) {
        RTB_CUDA(cudaMemsetAsync(fuse->acc, 0, sizeof(unsigned long long) * (size_t)(Us + 1), stream));
        int rc = accum_sum_rows(fuse->values, fuse->vdt, ikey_out, RTB_I32, nrows, Us + 1, fuse->bin_low, fuse->nanskip, fuse->acc, stream);
        if (rc != RTB_OK) return rc;
        RTB_CUDA(cudaStreamSynchronize(stream));
      } else if (fuse) {
        RTB_CUDA(cudaMemsetAsync(fuse->acc, 0, sizeof(unsigned long long) * (size_t)(Us + 1), stream));
      }
      st->cap = want_cap;
      st->unique = Us;
      st->prev_new = Us;
      st->prev_rows = nrows;
      st->rounds = 1;
      st->max_unique = max_unique;
      *unique_count_host = Us;
      return RTB_OK;
    }
  }

  uint32_t snap_cap = 0;   // inline-key snapshot (16-byte packed rows): entries in use, bins it holds
  int64_t snap_built = 0;
  uint32_t cap = st->cap;
  int64_t U = st->unique;
  int64_t prev_new = st->prev_new, prev_rows = st->prev_rows;
  int64_t lo = 0;
  int round = st->rounds;  // rounds since the table was created (0 = nothing known yet)
  int cround = 0;          // rounds of this chunk: every chunk restarts the geometric schedule with a small probe round
  while (lo < nrows) {
    // a chunk that resumes a settled table (nothing new in the previous chunk's last round) does not probe with a small
    // round again: it starts with the largest round a settled table may take
    const bool resumed_settled = cround == 0 && round > 0 && (double)prev_new <= plan.discovery_frac * (double)prev_rows;
    int64_t hi = cround == 0 ? (resumed_settled ? plan.finish_rows : plan.round0) : lo * plan.ratio;
    // a settled table and a short remainder: one round to the end saves the host round trips of the schedule (a burst of
    // late keys then costs at most one repeat of this short round)
    if (cround > 0 && nrows - lo <= plan.finish_rows && (double)prev_new <= plan.discovery_frac * (double)prev_rows) hi = nrows;
    if (hi > nrows || hi <= lo) hi = nrows;
    int64_t rows = hi - lo;
    // projected new keys in this round: discovery rate of the previous round, never above the row count
    double proj = round == 0 ? (double)rows : (double)prev_new * ((double)rows / (double)prev_rows) * 1.25 + 64;
    if (proj > (double)rows) proj = (double)rows;
    if (proj > (double)(max_unique - U)) proj = (double)(max_unique - U);
    uint32_t ideal = next_pow2_u64((uint64_t)ceil(((double)U + proj) / plan.load_target));
    // 4 Mi slots (64 MB) is the largest table whose random probes still run at the L2 rate (a table probed from both dies
    // can count on about half of the nominal L2: 2 M keys probe at 65 G rows/s in 128 MB and at 87 G rows/s in 64 MB at
    // twice the load), so a single-key table stays there while its load allows
    // ... and beyond it the smaller of two tables wins again (3 M keys: 48 G rows/s in 256 MB at 19 % load), so a table
    // outside the window is sized for twice the usual load
    if ((direct || plan.window_packed) && ideal > plan.window_slots) {
      if ((double)U + proj <= (double)plan.window_slots * plan.window_load) ideal = plan.window_slots;
      else if (ideal <= (1u << 25)) {  // (no gain measured above 512 MB: 10 M keys probe at 22 G rows/s either way)
        const uint32_t dense = next_pow2_u64((uint64_t)ceil(((double)U + proj) / plan.dense_load));
        if (dense >= plan.window_slots && dense < ideal) ideal = dense;
      }
    }
    if (ideal < 1024) ideal = 1024;
    if (ideal > capmax) ideal = capmax;
    // Resizing re-stages every entry, so keep the table while it still fits: grow only when the projection would
    // cross the abort load, shrink only once discovery has died down and the table is at least twice the ideal size.
    uint32_t want = ideal;
    if (cap != 0) {
      const bool fits = (double)U + proj <= (double)cap * plan.load_abort;
      const bool settled = (double)prev_new <= plan.discovery_frac * (double)prev_rows;
      if (fits && !(settled && (uint64_t)cap >= 2ull * ideal)) want = cap;
    }

    bool packed_this_round = false;
    bool fused_round = false;   // this round's kernel summed the rows whose bin was final at probe time
    bool redo_sums = false;     // a fused round aborted half-way: its partial sums are discarded below
    cudaEvent_t tev_pre = nullptr, tev_post0 = nullptr;
    if (trace) { cudaEventCreate(&tev_pre); cudaEventRecord(tev_pre, stream); }
    float trace_resize_ms = 0.f;
    for (int attempt = 0;; attempt++) {  // attempt loop: (re)size, run the round, grow on abort
      fused_round = false;
      if (want != cap) {
        uint32_t staged_n = 0;
        if (cap != 0 && U > 0) {
          RTB_CUDA(cudaMemsetAsync(L.ctr, 0, sizeof(GbCounters), stream));
          gb_stage_entries<<<grid_for((int64_t)cap + 1, 256, 1), 256, 0, stream>>>(L.table, cap + 1, L.staged, L.ctr);
          RTB_LAUNCH_CHECK();
          staged_n = (uint32_t)U;
        }
        RTB_CUDA(cudaMemsetAsync(L.table, 0xFF, sizeof(Slot) * ((size_t)want + 1), stream));
        if (staged_n) {
          if (direct)
            gb_reinsert_direct<<<grid_for(staged_n, 256, 1), 256, 0, stream>>>(L.table, want - 1, L.staged, staged_n);
          else
            switch (W) {
              case 16: gb_reinsert_packed<1><<<grid_for(staged_n, 256, 1), 256, 0, stream>>>(L.packed, L.table, want - 1, L.staged, staged_n, rowsz); break;
              case 32: gb_reinsert_packed<2><<<grid_for(staged_n, 256, 1), 256, 0, stream>>>(L.packed, L.table, want - 1, L.staged, staged_n, rowsz); break;
              case 48: gb_reinsert_packed<3><<<grid_for(staged_n, 256, 1), 256, 0, stream>>>(L.packed, L.table, want - 1, L.staged, staged_n, rowsz); break;
              case 64: gb_reinsert_packed<4><<<grid_for(staged_n, 256, 1), 256, 0, stream>>>(L.packed, L.table, want - 1, L.staged, staged_n, rowsz); break;
              default: gb_reinsert_byref<<<grid_for(staged_n, 256, 1), 256, 0, stream>>>(kc, L.table, want - 1, L.staged, staged_n);
            }
          RTB_LAUNCH_CHECK();
        }
        cap = want;
      }
      int64_t room = (int64_t)((double)cap * plan.load_abort) - U;
      if (room > max_unique - U) room = max_unique - U;
      if (room < 0) room = 0;
      uint32_t new_limit = (uint32_t)(room > 0xFFFFFFF0ll ? 0xFFFFFFF0ll : room);
      RTB_CUDA(cudaMemsetAsync(L.ctr, 0, sizeof(GbCounters), stream));
      uint32_t mask = cap - 1;
      if (plan.l2_window) set_table_l2_window(stream, L.table, sizeof(Slot) * ((size_t)cap + 1), true);
      cudaEvent_t tev0 = nullptr;
      if (trace) { cudaEventCreate(&tev0); cudaEventRecord(tev0, stream); }
      prof_begin(PROF_GB_ROUND, stream);
      if (direct) {
        // discovery rounds (many new keys expected) use the insert-oriented kernel, the rest the bucket-probe kernel
        // ... and so do rounds over a very large table: every probe is then a DRAM access, and at 17 GB (C5) the slot-wise
        // kernel's 16-byte probes run at 22 G rows/s where the 32-byte bucket probes reach 16.  (At 1-4 GB the two
        // kernels measure within 10 % of each other, 40-45 G rows/s.)
        const bool discovery = round == 0 || (double)prev_new > plan.discovery_frac * (double)prev_rows || attempt > 0 ||
                               (double)cap * sizeof(Slot) > plan.big_table_bytes;
        if (discovery) {
          int g2 = grid_for(rows, 256, 4, 8);
          switch (dkey.itemsize) {
            case 8: gb_round_insert<8><<<g2, 256, 0, stream>>>(dkey.data, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit); break;
            case 4: gb_round_insert<4><<<g2, 256, 0, stream>>>(dkey.data, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit); break;
            case 2: gb_round_insert<2><<<g2, 256, 0, stream>>>(dkey.data, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit); break;
            default: gb_round_insert<1><<<g2, 256, 0, stream>>>(dkey.data, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit);
          }
        } else if (cap <= kSmemCapMax && plan.smem_table) {
          // a full table in shared memory can loop: every slot final and none matching.  The table is at most 75 % full.
          const int isz = dkey.itemsize;
          const size_t smem = (size_t)cap * ((isz == 8 ? 8 : 4) + 4);
          int g3 = (int)((rows + 128 * 32 - 1) / (128 * 32));
          if (g3 > kSMs) g3 = kSMs;
          if (g3 < 1) g3 = 1;
#define RTB_SMEM_ROUND(ISZ_)                                                                                              \
  do {                                                                                                                    \
    static bool attr = false;                                                                                             \
    if (!attr) {                                                                                                          \
      RTB_CUDA(cudaFuncSetAttribute(gb_round_smem<ISZ_>, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                                    (int)(kSmemCapMax * ((ISZ_ == 8 ? 8 : 4) + 4))));                                     \
      attr = true;                                                                                                        \
    }                                                                                                                     \
    gb_round_smem<ISZ_><<<g3, 1024, smem, stream>>>(dkey.data, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr,  \
                                                   new_limit);                                                            \
  } while (0)
          switch (isz) {
            case 8: RTB_SMEM_ROUND(8); break;
            case 4: RTB_SMEM_ROUND(4); break;
            case 2: RTB_SMEM_ROUND(2); break;
            default: RTB_SMEM_ROUND(1);
          }
#undef RTB_SMEM_ROUND
        } else {
        fused_round = fuse_kernel_ok;
        const GbFuse fz = fused_round ? *fuse : fz_none;
        // a fused round that runs out of table room has already added some of its rows: keep a copy of the live
        // accumulators (U + 1 of them; provisional rows are never added by the round kernel) to fall back on
        if (fused_round)
          RTB_CUDA(cudaMemcpyAsync(fuse->snap, fuse->acc, sizeof(unsigned long long) * (size_t)(U + 1), cudaMemcpyDeviceToDevice, stream));
        const int fk = !fused_round ? 0 : (fuse->vsz == 8 ? 1 : 2);
#define RTB_ROUND(ISZ_, PF_, FK_)                                                                                          \
  do {                                                                                                                    \
    static int occ = 0;                                                                                                   \
    if (!occ) {                                                                                                           \
      RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gb_round_direct<ISZ_, PF_, FK_>, 256, 0));             \
      if (occ < 1) occ = 1;                                                                                               \
    }                                                                                                                     \
    const int grid = grid_for(rows, 256, 4, occ < plan.ctas_per_sm ? occ : plan.ctas_per_sm);                             \
    gb_round_direct<ISZ_, PF_, FK_><<<grid, 256, 0, stream>>>(dkey.data, filter, ikey_out, L.table, mask, lo, hi, newlist, \
                                                              L.ctr, new_limit, fz);                                      \
  } while (0)
#define RTB_ROUND_F(ISZ_, PF_)                                                                                            \
  do {                                                                                                                    \
    if (fk == 1) RTB_ROUND(ISZ_, PF_, 1);                                                                                 \
    else if (fk == 2) RTB_ROUND(ISZ_, PF_, 2);                                                                            \
    else RTB_ROUND(ISZ_, PF_, 0);                                                                                         \
  } while (0)
        switch (dkey.itemsize) {
          case 8: if (fk ? plan.fuse_prefetch : plan.prefetch) RTB_ROUND_F(8, true); else RTB_ROUND_F(8, false); break;
          case 4: if (fk ? plan.fuse_prefetch : plan.prefetch) RTB_ROUND_F(4, true); else RTB_ROUND_F(4, false); break;
          case 2: RTB_ROUND_F(2, false); break;
          default: RTB_ROUND_F(1, false);
        }
#undef RTB_ROUND_F
#undef RTB_ROUND
        }
      } else if (W && !has_cut && round > 0 && (double)prev_new <= plan.discovery_frac * (double)prev_rows && attempt == 0 &&
                 plan.smem_table && U > 0 && U < 65535 &&
                 row_offset == 0 && packed_smem_bytes(U, W) <= kPackedSmemBytes) {
        // settled, and every key fits in shared memory: stream-bound round (ifirstkey holds the representative rows)
        const uint32_t icap = packed_smem_icap(U);
        const size_t smem = packed_smem_bytes(U, W);
        const int from_cols = (lazy_pack && W <= 32 && cols_aligned) ? (whole_words_all ? 2 : 1) : 0;
        if (lazy_pack && !from_cols && !packed_this_round) {
          int prc = pack_range(lo, hi);
          if (prc != RTB_OK) return prc;
          packed_this_round = true;
        }
        int g3 = (int)((rows + 1024 * 8 - 1) / (1024 * 8));
        if (g3 > kSMs) g3 = kSMs;
        if (g3 < 1) g3 = 1;
#define RTB_PSMEM_ROUND(W16_)                                                                                             \
  do {                                                                                                                    \
    static bool attr = false;                                                                                             \
    if (!attr) {                                                                                                          \
      RTB_CUDA(cudaFuncSetAttribute(gb_round_packed_smem<W16_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPackedSmemBytes)); \
      attr = true;                                                                                                        \
    }                                                                                                                     \
    gb_round_packed_smem<W16_><<<g3, 1024, smem, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, \
                                                          new_limit, ifirstkey_out, (uint32_t)U, icap, rowsz, kc, from_cols, wm); \
  } while (0)
        switch (W) {
          case 16: RTB_PSMEM_ROUND(1); break;
          case 32: RTB_PSMEM_ROUND(2); break;
          case 48: RTB_PSMEM_ROUND(3); break;
          default: RTB_PSMEM_ROUND(4);
        }
#undef RTB_PSMEM_ROUND
      } else if ((W == 16 || W == 32) && L.snap != nullptr && plan.snapshot && !has_cut && round > 0 && attempt == 0 && row_offset == 0 &&
                 U > 0 && (double)prev_new <= plan.discovery_frac * (double)prev_rows && 2 * U <= (int64_t)snap_cap_for(max_unique)) {
        // settled 16- / 32-byte rows beyond shared memory: one-access probes of the inline-key snapshot
        const size_t esz = (size_t)2 * W;
        if (snap_cap == 0 || 2 * U > (int64_t)snap_cap) {
          uint32_t c = 1024;
          while ((int64_t)c < (int64_t)(plan.snap_fill * (double)U)) c <<= 1;
          const uint32_t cmax = snap_cap_for(max_unique);
          snap_cap = c > cmax ? cmax : c;
          RTB_CUDA(cudaMemsetAsync(L.snap, 0, esz * (size_t)snap_cap, stream));
          snap_built = 0;
        }
        const int from_cols = (lazy_pack && cols_aligned) ? (whole_words_all ? 2 : 1) : 0;
        const int gs = grid_for(rows, 256, 4, 3);
        if (W == 16) {
          if (snap_built < U) {
            gb_snap_build<1><<<grid_for(U - snap_built, 256, 1), 256, 0, stream>>>(L.packed, rowsz, ifirstkey_out, (uint32_t)snap_built,
                                                                                   (uint32_t)U, (SnapEntry<1>*)L.snap, snap_cap - 1);
            RTB_LAUNCH_CHECK();
          }
          gb_round_packed_snap<1><<<gs, 256, 0, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit,
                                                         (const SnapEntry<1>*)L.snap, snap_cap - 1, rowsz, kc, from_cols, wm);
        } else {
          if (snap_built < U) {
            gb_snap_build<2><<<grid_for(U - snap_built, 256, 1), 256, 0, stream>>>(L.packed, rowsz, ifirstkey_out, (uint32_t)snap_built,
                                                                                   (uint32_t)U, (SnapEntry<2>*)L.snap, snap_cap - 1);
            RTB_LAUNCH_CHECK();
          }
          gb_round_packed_snap<2><<<gs, 256, 0, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit,
                                                         (const SnapEntry<2>*)L.snap, snap_cap - 1, rowsz, kc, from_cols, wm);
        }
        snap_built = U;
      } else {
        if (lazy_pack && !packed_this_round) {
          int prc = pack_range(lo, hi);
          if (prc != RTB_OK) return prc;
          packed_this_round = true;
        }
        // a table far beyond L2 and a round that mostly looks keys up: fast path first, everything else in a second launch
        const bool defer = plan.defer_packed && (W == 16 || W == 32 || W == 48 || W == 64) && round > 0 && attempt == 0 && cap <= (1u << 30) &&
                           (double)cap * sizeof(Slot) > plan.defer_table_bytes && (double)prev_new <= 0.5 * (double)prev_rows;
        if (defer) {
          const int gf = grid_for(rows, 256, 2, 8), gd = grid_for(rows, 256, 8, 8);
#define RTB_DEFER_ROUND(W16_)                                                                                               \
  do {                                                                                                                    \
    gb_round_packed_fast<W16_><<<gf, 256, 0, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, rowsz);          \
    RTB_LAUNCH_CHECK();                                                                                                   \
    gb_round_packed_deferred<W16_><<<gd, 256, 0, stream>>>(L.packed, ikey_out, L.table, mask, lo, hi, newlist, L.ctr,      \
                                                           new_limit, rowsz);                                             \
  } while (0)
          switch (W) {
            case 16: RTB_DEFER_ROUND(1); break;
            case 32: RTB_DEFER_ROUND(2); break;
            case 48: RTB_DEFER_ROUND(3); break;
            default: RTB_DEFER_ROUND(4);
          }
#undef RTB_DEFER_ROUND
        } else
        switch (W) {
          case 16: gb_round_packed<1><<<grid_for(rows, 256, 2, 8), 256, 0, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit, rowsz); break;
          case 32: gb_round_packed<2><<<grid_for(rows, 256, 2, 8), 256, 0, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit, rowsz); break;
          case 48: gb_round_packed<3><<<grid_for(rows, 256, 2, 8), 256, 0, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit, rowsz); break;
          case 64: gb_round_packed<4><<<grid_for(rows, 256, 2, 8), 256, 0, stream>>>(L.packed, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit, rowsz); break;
          default: gb_round_byref<<<grid_for(rows, 256, 4, 6), 256, 0, stream>>>(kc, filter, ikey_out, L.table, mask, lo, hi, newlist, L.ctr, new_limit);
        }
      }
      RTB_LAUNCH_CHECK();
      if (plan.l2_window) set_table_l2_window(stream, nullptr, 0, false);
      prof_end(PROF_GB_ROUND, stream, (uint64_t)rows);
      cudaEvent_t tev1 = nullptr;
      if (trace) { cudaEventCreate(&tev1); cudaEventRecord(tev1, stream); }
      // the round's counters come back through a page-locked buffer (a pageable target costs the driver a staging copy
      // on every round)
      static thread_local GbCounters* pinned_ctr = nullptr;
      if (!pinned_ctr && cudaHostAlloc((void**)&pinned_ctr, sizeof(GbCounters), cudaHostAllocDefault) != cudaSuccess) {
        pinned_ctr = nullptr;
        cudaGetLastError();
      }
      GbCounters h_local;
      GbCounters* hp = pinned_ctr ? pinned_ctr : &h_local;
      RTB_CUDA(cudaMemcpyAsync(hp, L.ctr, sizeof(GbCounters), cudaMemcpyDeviceToHost, stream));
      RTB_CUDA(cudaStreamSynchronize(stream));
      const GbCounters h = *hp;
      if (trace) {
        float tms = 0.f;
        cudaEventElapsedTime(&tms, tev0, tev1);
        cudaEventElapsedTime(&trace_resize_ms, tev_pre, tev0);
        fprintf(stderr, "[rtb gb] round %d rows %lld cap %u U %lld new %u abort %u resize %.3f ms kernel %.3f ms (%.1f G rows/s)\n", round,
                (long long)rows, cap, (long long)U, h.newcount, h.abort, trace_resize_ms, tms, rows / (tms * 1e6));
        cudaEventDestroy(tev0);
        cudaEventDestroy(tev1);
      }
      if (!h.abort) {
        prev_new = h.newcount;
        break;
      }
      if (fused_round) redo_sums = true;
      // out of room: grow 4x (provisional entries are dropped by the re-stage) or ask for more workspace
      if (cap >= capmax) {
        double seen = (double)(lo > 0 ? lo : 1);
        double est = (double)(U + room + 1) * ((double)nrows / seen);
        if (est > (double)nrows) est = (double)nrows;
        if (est < 4.0 * (double)max_unique) est = 4.0 * (double)max_unique;
        if (est > (double)nrows) est = (double)nrows;
        if (needed_unique_host) *needed_unique_host = (int64_t)est;
        set_error("MultiKeyGroupBy32: more than %lld unique keys; retry with max_unique >= %lld", (long long)max_unique,
                  (long long)est);
        return RTB_ERR_WORKSPACE;
      }
      uint64_t g = (uint64_t)cap * 4;
      want = g > capmax ? capmax : (uint32_t)g;
    }

    if (trace) { cudaEventCreate(&tev_post0); cudaEventRecord(tev_post0, stream); }
    if (fuse_values && redo_sums)  // the aborted attempt's partial sums go; the repeat ran unfused and is added below
      RTB_CUDA(cudaMemcpyAsync(fuse->acc, fuse->snap, sizeof(unsigned long long) * (size_t)(U + 1), cudaMemcpyDeviceToDevice, stream));
    if (fuse && prev_new > 0)
      RTB_CUDA(cudaMemsetAsync(fuse->acc + U + 1, 0, sizeof(unsigned long long) * (size_t)prev_new, stream));
    if (prev_new > 0) {
      int64_t words = (rows + 31) / 32;
      RTB_ARG(words <= L.bitmap_words, "MultiKeyGroupBy32: internal error, a round of %lld rows exceeds the first-row bitmap", (long long)rows);
      RTB_CUDA(cudaMemsetAsync(L.bitmap, 0, sizeof(uint32_t) * words, stream));
      uint32_t* minrows = newlist + ((size_t)max_unique + 1);  // the list uses at most max_unique of the staging area's 4 (max_unique + 1) words
      gb_mark_firstrows<<<grid_for(prev_new, 256, 1), 256, 0, stream>>>(L.table, newlist, L.ctr, lo, L.bitmap, minrows);
      RTB_LAUNCH_CHECK();
      int rc = exclusive_scan_u32(L.bitmap, L.prefix, words, SCAN_POPC, L.scan_scratch, nullptr, stream);
      if (rc != RTB_OK) return rc;
      gb_assign_bins<<<grid_for(prev_new, 256, 1), 256, 0, stream>>>(L.table, newlist, L.ctr, L.bitmap, L.prefix, (uint32_t)U, minrows);
      RTB_LAUNCH_CHECK();
      gb_firstkeys_from_bitmap<<<grid_for(words, 256, 1), 256, 0, stream>>>(L.bitmap, L.prefix, words, lo, (uint32_t)U, ifirstkey_out,
                                                                           -row_offset);
      RTB_LAUNCH_CHECK();
      gb_fixup<<<grid_for(rows, 256, 4), 256, 0, stream>>>(ikey_out, L.table, lo, hi, 0, fused_round ? *fuse : fz_none);
      RTB_LAUNCH_CHECK();
      U += prev_new;
    }
    if (fuse_values && !fused_round) {
      // rounds whose kernel has no fused form (discovery, shared-memory tables, multi-key rows): the round's rows are
      // summed from the iKey it just wrote, while that is still in L2 for all but the largest rounds
      const int vsz = fuse->vsz;
      int rc = accum_sum_rows((const uint8_t*)fuse->values + lo * vsz, fuse->vdt, ikey_out + lo, RTB_I32, rows, U + 1,
                              fuse->bin_low, fuse->nanskip, fuse->acc, stream);
      if (rc != RTB_OK) return rc;
    }
    if (trace) {
      cudaEvent_t tev_post1;
      cudaEventCreate(&tev_post1);
      cudaEventRecord(tev_post1, stream);
      cudaEventSynchronize(tev_post1);
      float pms = 0.f;
      cudaEventElapsedTime(&pms, tev_post0, tev_post1);
      fprintf(stderr, "[rtb gb]   first rows / bins / fix-up%s: %.3f ms\n", (fuse_values && !fused_round) ? " / sums" : "", pms);
      cudaEventDestroy(tev_pre);
      cudaEventDestroy(tev_post0);
      cudaEventDestroy(tev_post1);
    }
    prev_rows = rows;
    lo = hi;
    round++;
    cround++;
  }
  st->cap = cap;
  st->unique = U;
  st->prev_new = prev_new;
  st->prev_rows = prev_rows;
  st->rounds = round;
  st->max_unique = max_unique;

  if (has_cut) {
    // global bins -> numbering that restarts per partition; iFirstKey -> partition-relative rows.
    // Bins are in first-appearance order and partitions are contiguous row ranges, so partition p
    // owns the contiguous bin range [ubase[p], ubase[p+1]).
    RTB_CUDA(cudaMemsetAsync(L.ucount, 0, sizeof(uint32_t) * ncutoffs, stream));
    if (U > 0) {
      gb_count_partition_uniques<<<grid_for(U, 256, 1), 256, 0, stream>>>(ifirstkey_out, U, L.pid, L.ucount);
      RTB_LAUNCH_CHECK();
    }
    uint32_t* hc = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)ncutoffs);
    if (!hc) { set_error("out of host memory"); return RTB_ERR_ARG; }
    cudaError_t e = cudaMemcpyAsync(hc, L.ucount, sizeof(uint32_t) * ncutoffs, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    uint32_t run = 0;
    for (int64_t p = 0; p < ncutoffs && e == cudaSuccess; p++) {
      uint32_t c = hc[p];
      hc[p] = run;
      run += c;
      if (ucutoffs_host) ucutoffs_host[p] = run;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(L.ucount, hc, sizeof(uint32_t) * ncutoffs, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    free(hc);
    RTB_CUDA(e);
    if (U > 0) {
      gb_relative_firstkeys<<<grid_for(U, 256, 1), 256, 0, stream>>>(ifirstkey_out, U, L.pid, L.cutoffs_dev);
      RTB_LAUNCH_CHECK();
    }
    gb_renumber_partitions<<<grid_for(nrows, 256, 1), 256, 0, stream>>>(ikey_out, nrows, L.pid, L.ucount);
    RTB_LAUNCH_CHECK();
  }

  RTB_CUDA(cudaStreamSynchronize(stream));
  *unique_count_host = U;
  return RTB_OK;
}

extern "C" int rtb_multikey_groupby32(const rtb_column* keys, int nkeys, int64_t nrows, int64_t hint_size,
                                      const uint8_t* filter, const int64_t* cutoffs_host, int64_t ncutoffs,
                                      int32_t* ikey_out, int32_t* ifirstkey_out, int64_t max_unique,
                                      int64_t* unique_count_host, int64_t* ucutoffs_host, int64_t* needed_unique_host,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
  (void)hint_size;  // the shim folds hint_size into max_unique; the table itself is sized adaptively
  RTB_ARG(keys && nkeys >= 1, "MultiKeyGroupBy32: first argument is not a list of arrays");
  RTB_ARG(nkeys <= kMaxKeyCols - 1, "MultiKeyGroupBy32: at most %d key columns", kMaxKeyCols - 1);
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll, "MultiKeyGroupBy32: row count out of range for int32 iKey");
  RTB_ARG(unique_count_host, "MultiKeyGroupBy32: null unique_count");
  *unique_count_host = 0;
  if (needed_unique_host) *needed_unique_host = 0;
  if (nrows == 0) return RTB_OK;
  RTB_ARG(ikey_out && ifirstkey_out && max_unique >= 1, "MultiKeyGroupBy32: null output");
  RTB_ARG(((uintptr_t)ikey_out % 16) == 0, "MultiKeyGroupBy32: iKey buffer must be 16-byte aligned");
  for (int c = 0; c < nkeys; c++) RTB_ARG(keys[c].data && keys[c].itemsize >= 1, "MultiKeyGroupBy32: bad key column %d", c);
  if (cutoffs_host != nullptr && ncutoffs > 0) {
    RTB_ARG(cutoffs_host[ncutoffs - 1] == nrows, "MultiKeyGroupBy32: last cutoff must equal the row count");
    for (int64_t i = 1; i < ncutoffs; i++)
      RTB_ARG(cutoffs_host[i] >= cutoffs_host[i - 1], "MultiKeyGroupBy32: cutoffs must be non-decreasing");
  }
  if (max_unique > nrows) max_unique = nrows;
  rtb_gb_state st;
  memset(&st, 0, sizeof(st));
  return gb_core(keys, nkeys, nrows, filter, cutoffs_host, ncutoffs, ikey_out, ifirstkey_out, max_unique, unique_count_host,
                 ucutoffs_host, needed_unique_host, workspace, workspace_bytes, (cudaStream_t)stream_, &st, 0);
}

extern "C" int rtb_multikey_groupby32_chunk(const rtb_column* key, int64_t nrows, const uint8_t* filter, int32_t* ikey_out,
                                            int32_t* ifirstkey_out, int64_t max_unique, int64_t row_offset,
                                            rtb_gb_state* state, int resume, int64_t* needed_unique_host, void* workspace,
                                            size_t workspace_bytes, void* stream_) {
  RTB_ARG(key && state, "groupby chunk: null key or state");
  RTB_ARG(key->itemsize == 1 || key->itemsize == 2 || key->itemsize == 4 || key->itemsize == 8,
          "groupby chunk: one key column of 1, 2, 4 or 8 bytes");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll, "groupby chunk: row count out of range");
  RTB_ARG(max_unique >= 1 && ifirstkey_out, "groupby chunk: bad max_unique or null iFirstKey");
  RTB_ARG(row_offset >= 0 && row_offset + nrows <= 2147483646ll, "groupby chunk: first rows must fit int32");
  if (!resume) memset(state, 0, sizeof(*state));
  else RTB_ARG(state->max_unique == max_unique, "groupby chunk: max_unique must not change between chunks");
  state->max_unique = max_unique;
  if (needed_unique_host) *needed_unique_host = 0;
  if (nrows == 0) return RTB_OK;
  RTB_ARG(key->data && ikey_out && ((uintptr_t)ikey_out % 16) == 0 && ((uintptr_t)key->data % 16) == 0 &&
              (!filter || ((uintptr_t)filter % 4) == 0),
          "groupby chunk: buffers must be 16-byte aligned");
  int64_t u = 0;
  return gb_core(key, 1, nrows, filter, nullptr, 0, ikey_out, ifirstkey_out, max_unique, &u, nullptr, needed_unique_host,
                 workspace, workspace_bytes, (cudaStream_t)stream_, state, row_offset);
}

// ---- fused grouping + per-bin sum ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gb_cast_sums_f32(const unsigned long long* __restrict__ acc, float* __restrict__ out, int64_t nb) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x)
    out[b] = (float)__longlong_as_double((long long)acc[b]);
}

extern "C" size_t rtb_groupby_sum_workspace_bytes(int64_t nrows, int64_t max_unique, int64_t row_bytes) {
  if (nrows < 0 || max_unique < 0 || row_bytes < 0) return 0;
  return align_up(rtb_groupby_workspace_bytes_rows(nrows, max_unique, row_bytes, 0), 256) +
         2 * align_up(sizeof(unsigned long long) * ((size_t)max_unique + 1), 256) + 256;
}

extern "C" int rtb_multikey_groupby_sum32(const rtb_column* keys, int nkeys, int64_t nrows, const uint8_t* filter,
                                          const void* values, int vdtype, int func, int64_t bin_low, int32_t* ikey_out,
                                          int32_t* ifirstkey_out, int64_t max_unique, void* sum_out,
                                          int64_t* unique_count_host, int64_t* needed_unique_host, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(keys && nkeys >= 1, "MultiKeyGroupBySum32: first argument is not a list of arrays");
  RTB_ARG(nkeys <= kMaxKeyCols - 1, "MultiKeyGroupBySum32: at most %d key columns", kMaxKeyCols - 1);
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll, "MultiKeyGroupBySum32: row count out of range for int32 iKey");
  RTB_ARG(unique_count_host, "MultiKeyGroupBySum32: null unique_count");
  RTB_ARG(func == RTB_GB_SUM || func == RTB_GB_NANSUM, "MultiKeyGroupBySum32: GB_SUM or GB_NANSUM");
  RTB_ARG(bin_low == 0 || bin_low == 1, "MultiKeyGroupBySum32: bin_low must be 0 or 1");
  const int odt = rtb_reduce_out_dtype(func, vdtype);
  if (odt < 0) {
    set_error("MultiKeyGroupBySum32: dtype %d cannot be summed", vdtype);
    return RTB_ERR_UNSUPPORTED;
  }
  *unique_count_host = 0;
  if (needed_unique_host) *needed_unique_host = 0;
  RTB_ARG(sum_out, "MultiKeyGroupBySum32: null output");
  if (nrows == 0) {
    RTB_CUDA(cudaMemsetAsync(sum_out, 0, dtype_size(odt), stream));
    return RTB_OK;
  }
  RTB_ARG(ikey_out && ifirstkey_out && max_unique >= 1 && values, "MultiKeyGroupBySum32: null pointer");
  RTB_ARG(((uintptr_t)ikey_out % 16) == 0 && ((uintptr_t)values % 16) == 0 && ((uintptr_t)sum_out % 8) == 0,
          "MultiKeyGroupBySum32: iKey and value buffers must be 16-byte aligned");
  for (int c = 0; c < nkeys; c++) RTB_ARG(keys[c].data && keys[c].itemsize >= 1, "MultiKeyGroupBySum32: bad key column %d", c);
  if (max_unique > nrows) max_unique = nrows;
  int64_t row_bytes = 0;
  for (int c = 0; c < nkeys; c++) row_bytes += keys[c].itemsize;
  const bool single = nkeys == 1 && (keys[0].itemsize == 1 || keys[0].itemsize == 2 || keys[0].itemsize == 4 || keys[0].itemsize == 8);
  const size_t gb_bytes = align_up(rtb_groupby_workspace_bytes_rows(nrows, max_unique, single ? 0 : row_bytes, 0), 256);
  const size_t acc_bytes = align_up(sizeof(unsigned long long) * ((size_t)max_unique + 1), 256);
  RTB_ARG(workspace && ((uintptr_t)workspace % 256) == 0 && gb_bytes + 2 * acc_bytes <= workspace_bytes,
          "MultiKeyGroupBySum32: workspace too small or not 256-byte aligned");
  GbFuse fz;
  memset(&fz, 0, sizeof(fz));
  fz.values = values;
  fz.vdt = vdtype;
  fz.vsz = dtype_size(vdtype);
  fz.nanskip = func == RTB_GB_NANSUM;
  fz.isfloat = vdtype == RTB_F32 || vdtype == RTB_F64;
  fz.bin_low = (int)bin_low;
  // 64-bit outputs accumulate in place; float32 sums accumulate in f64 and are rounded once at the end
  fz.acc = vdtype == RTB_F32 ? (unsigned long long*)((uint8_t*)workspace + gb_bytes) : (unsigned long long*)sum_out;
  fz.snap = (unsigned long long*)((uint8_t*)workspace + gb_bytes + acc_bytes);
  rtb_gb_state st;
  memset(&st, 0, sizeof(st));
  int rc = gb_core(keys, nkeys, nrows, filter, nullptr, 0, ikey_out, ifirstkey_out, max_unique, unique_count_host, nullptr,
                   needed_unique_host, workspace, gb_bytes, stream, &st, 0, &fz);
  if (rc != RTB_OK) return rc;
  if (vdtype == RTB_F32) {
    const int64_t nb = *unique_count_host + 1;
    gb_cast_sums_f32<<<grid_for(nb, 256, 1), 256, 0, stream>>>(fz.acc, (float*)sum_out, nb);
    RTB_LAUNCH_CHECK();
  }
  return RTB_OK;
}

extern "C" size_t rtb_groupby_sum_chunk_workspace_bytes(int64_t max_chunk_rows, int64_t max_unique) {
  if (max_chunk_rows < 0 || max_unique < 0) return 0;
  return align_up(rtb_groupby_workspace_bytes(max_chunk_rows, max_unique, 1, 0), 256) +
         align_up(sizeof(unsigned long long) * ((size_t)max_unique + 1), 256) + 256;
}

extern "C" int rtb_multikey_groupby_sum32_chunk(const rtb_column* key, int64_t nrows, const uint8_t* filter,
                                                const void* values, int vdtype, int func, int64_t bin_low,
                                                int32_t* ikey_out, int32_t* ifirstkey_out, int64_t max_unique,
                                                void* acc, int64_t row_offset, rtb_gb_state* state, int resume,
                                                int64_t* needed_unique_host, void* workspace, size_t workspace_bytes,
                                                void* stream_) {
  RTB_ARG(key && state && acc, "groupby-sum chunk: null key, state or accumulators");
  RTB_ARG(key->itemsize == 1 || key->itemsize == 2 || key->itemsize == 4 || key->itemsize == 8,
          "groupby-sum chunk: one key column of 1, 2, 4 or 8 bytes");
  RTB_ARG(nrows >= 0 && nrows <= 2147483646ll, "groupby-sum chunk: row count out of range");
  RTB_ARG(max_unique >= 1 && ifirstkey_out, "groupby-sum chunk: bad max_unique or null iFirstKey");
  RTB_ARG(row_offset >= 0 && row_offset + nrows <= 2147483646ll, "groupby-sum chunk: first rows must fit int32");
  RTB_ARG(func == RTB_GB_SUM || func == RTB_GB_NANSUM, "groupby-sum chunk: GB_SUM or GB_NANSUM");
  RTB_ARG(bin_low == 0 || bin_low == 1, "groupby-sum chunk: bin_low must be 0 or 1");
  RTB_ARG(!values || rtb_reduce_out_dtype(func, vdtype) >= 0, "groupby-sum chunk: the value dtype cannot be summed");
  if (!resume) memset(state, 0, sizeof(*state));
  else RTB_ARG(state->max_unique == max_unique, "groupby-sum chunk: max_unique must not change between chunks");
  state->max_unique = max_unique;
  if (needed_unique_host) *needed_unique_host = 0;
  if (nrows == 0) {
    if (!resume) RTB_CUDA(cudaMemsetAsync(acc, 0, sizeof(unsigned long long), (cudaStream_t)stream_));
    return RTB_OK;
  }
  RTB_ARG(key->data && ikey_out && ((uintptr_t)ikey_out % 16) == 0 && ((uintptr_t)key->data % 16) == 0 &&
              (!filter || ((uintptr_t)filter % 4) == 0) && ((uintptr_t)values % 16) == 0 && ((uintptr_t)acc % 8) == 0,
          "groupby-sum chunk: buffers must be 16-byte aligned");
  const size_t gb_bytes = align_up(rtb_groupby_workspace_bytes(nrows, max_unique, 1, 0), 256);
  const size_t acc_bytes = align_up(sizeof(unsigned long long) * ((size_t)max_unique + 1), 256);
  RTB_ARG(workspace && ((uintptr_t)workspace % 256) == 0 && workspace_bytes >= acc_bytes + gb_bytes,
          "groupby-sum chunk: workspace too small or not 256-byte aligned");
  GbFuse fz;
  memset(&fz, 0, sizeof(fz));
  fz.values = values;
  fz.vdt = vdtype;
  fz.vsz = values ? dtype_size(vdtype) : 0;
  fz.nanskip = func == RTB_GB_NANSUM;
  fz.isfloat = vdtype == RTB_F32 || vdtype == RTB_F64;
  fz.bin_low = (int)bin_low;
  fz.acc = (unsigned long long*)acc;
  // the hash table sits at the start of the workspace and must stay put between chunks; the snapshot goes behind the
  // grouping scratch of the LARGEST chunk (the size the caller provisioned), i.e. at the very end
  fz.snap = (unsigned long long*)((uint8_t*)workspace + ((workspace_bytes - acc_bytes) & ~(size_t)255));
  int64_t u = 0;
  return gb_core(key, 1, nrows, filter, nullptr, 0, ikey_out, ifirstkey_out, max_unique, &u, nullptr, needed_unique_host,
                 workspace, (size_t)((uint8_t*)fz.snap - (uint8_t*)workspace), (cudaStream_t)stream_, state, row_offset, &fz);
}

extern "C" int rtb_multikey_groupby32_adopt_seed(const rtb_column* seed_keys, int64_t nseed, int32_t* ifirstkey,
                                                 int64_t max_unique, void* acc, rtb_gb_state* state, int32_t* map_out,
                                                 int64_t* nlate_host, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(seed_keys && state && map_out && nlate_host && ifirstkey, "adopt seed: null argument");
  RTB_ARG(seed_keys->itemsize == 1 || seed_keys->itemsize == 2 || seed_keys->itemsize == 4 || seed_keys->itemsize == 8,
          "adopt seed: one key column of 1, 2, 4 or 8 bytes");
  RTB_ARG(nseed >= 0 && nseed <= 2147483646ll && state->max_unique == max_unique && state->cap != 0, "adopt seed: bad sizes or a fresh state");
  RTB_ARG(nseed == 0 || seed_keys->data, "adopt seed: null seed");
  *nlate_host = 0;
  const int64_t u_loc = state->unique;
  const uint32_t cap = state->cap;
  GbPlan plan = get_plan();
  // every seed key may be new to the table: it must have the room, and the bins must fit the provisioned arrays
  if ((double)(u_loc + nseed) > (double)cap * plan.load_abort || u_loc + nseed > max_unique) {
    set_error("adopt seed: the table has no room for %lld more keys", (long long)nseed);
    return RTB_ERR_UNSUPPORTED;
  }
  GbLayout L = layout(workspace, workspace_bytes, 1, max_unique, false, 0);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "adopt seed: workspace too small");
  // scratch inside the staging area (16 bytes per provisioned bin): [flags | excl | scan scratch], later the permuted copies
  uint8_t* scratch = (uint8_t*)L.staged;
  const size_t scratch_bytes = sizeof(Slot) * ((size_t)max_unique + 1);
  uint32_t* flags = (uint32_t*)scratch;
  uint32_t* excl = flags + u_loc + 4;
  uint32_t* scan_scratch = excl + u_loc + 4;
  const size_t scan_elems = scan_u32_workspace_elems(u_loc > 0 ? u_loc : 1) + 8;
  RTB_ARG((size_t)((uint8_t*)(scan_scratch + scan_elems) - scratch) <= scratch_bytes, "adopt seed: workspace too small for the scan");
  RTB_CUDA(cudaMemsetAsync(map_out, 0, sizeof(int32_t) * (size_t)(u_loc + 1), stream));
  if (nseed > 0) {
    const int g = grid_for(nseed, 256, 1);
    switch (seed_keys->itemsize) {
      case 8: gb_adopt_probe_kernel<8><<<g, 256, 0, stream>>>(seed_keys->data, nseed, L.table, cap - 1, map_out); break;
      case 4: gb_adopt_probe_kernel<4><<<g, 256, 0, stream>>>(seed_keys->data, nseed, L.table, cap - 1, map_out); break;
      case 2: gb_adopt_probe_kernel<2><<<g, 256, 0, stream>>>(seed_keys->data, nseed, L.table, cap - 1, map_out); break;
      default: gb_adopt_probe_kernel<1><<<g, 256, 0, stream>>>(seed_keys->data, nseed, L.table, cap - 1, map_out);
    }
    RTB_LAUNCH_CHECK();
  }
  uint32_t nlate = 0;
  if (u_loc > 0) {
    uint32_t* total_dev = scan_scratch + scan_elems - 4;
    gb_adopt_late_flags_kernel<<<grid_for(u_loc, 256, 1), 256, 0, stream>>>(map_out, u_loc, flags);
    RTB_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(flags, excl, u_loc, SCAN_IDENTITY, scan_scratch, total_dev, stream);
    if (rc != RTB_OK) return rc;
    gb_adopt_late_bins_kernel<<<grid_for(u_loc, 256, 1), 256, 0, stream>>>(map_out, u_loc, flags, excl, (uint32_t)nseed);
    RTB_LAUNCH_CHECK();
    gb_adopt_relabel_kernel<<<grid_for((int64_t)cap + 1, 256, 1), 256, 0, stream>>>(L.table, cap + 1, map_out);
    RTB_LAUNCH_CHECK();
    RTB_CUDA(cudaMemcpyAsync(&nlate, total_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    RTB_CUDA(cudaStreamSynchronize(stream));
  }
  const int64_t u_new = nseed + (int64_t)nlate;
  // iFirstKey: seed bins this shard has not seen get -1, the others keep their local first rows under the new numbering
  {
    int32_t* tmp = (int32_t*)scratch;
    RTB_CUDA(cudaMemsetAsync(tmp, 0xFF, sizeof(int32_t) * (size_t)(u_new > 0 ? u_new : 1), stream));
    if (u_loc > 0) {
      gb_adopt_permute_kernel<int32_t><<<grid_for(u_loc, 256, 1), 256, 0, stream>>>(ifirstkey, tmp, map_out, 1, u_loc, 1);
      RTB_LAUNCH_CHECK();
    }
    RTB_CUDA(cudaMemcpyAsync(ifirstkey, tmp, sizeof(int32_t) * (size_t)u_new, cudaMemcpyDeviceToDevice, stream));
  }
  if (acc) {
    unsigned long long* tmp = (unsigned long long*)scratch;  // stream order: the iFirstKey copy above is done by now
    RTB_CUDA(cudaMemsetAsync(tmp, 0, sizeof(unsigned long long) * (size_t)(u_new + 1), stream));
    gb_adopt_permute_kernel<unsigned long long><<<grid_for(u_loc + 1, 256, 1), 256, 0, stream>>>((const unsigned long long*)acc, tmp,
                                                                                                   map_out, 0, u_loc, 0);
    RTB_LAUNCH_CHECK();
    RTB_CUDA(cudaMemcpyAsync(acc, tmp, sizeof(unsigned long long) * (size_t)(u_new + 1), cudaMemcpyDeviceToDevice, stream));
  }
  RTB_CUDA(cudaStreamSynchronize(stream));
  state->unique = u_new;
  state->prev_new = 0;  // the table is settled: the next chunk starts with a large round
  state->prev_rows = u_new > 0 ? u_new : 1;
  *nlate_host = (int64_t)nlate;
  return RTB_OK;
}

struct IsMemberLayout {
  void* gb_ws;
  size_t gb_bytes;
  int32_t *ikey_b, *ifirst_b;
  uint8_t* valid_b;
  size_t bytes;
};
static IsMemberLayout ismember_layout(void* ws, size_t wsbytes, int64_t nb) {
  Arena a(ws, wsbytes);
  IsMemberLayout L;
  const int64_t m = nb > 0 ? nb : 1;
  L.gb_bytes = layout(nullptr, 0, m, m, false, 0).bytes + 256;
  L.gb_ws = a.take<uint8_t>(L.gb_bytes);
  L.ikey_b = a.take<int32_t>((size_t)m + 4);
  L.ifirst_b = a.take<int32_t>((size_t)m + 4);
  L.valid_b = a.take<uint8_t>((size_t)m + 4);
  L.bytes = a.off;
  return L;
}

extern "C" size_t rtb_ismember_workspace_bytes(int64_t nb) {
  if (nb < 0) return 0;
  return ismember_layout(nullptr, 0, nb).bytes + 256;
}

extern "C" int rtb_ismember32(const void* a, int64_t na, const void* b, int64_t nb, int dtype, int base_index, int index_dtype,
                              uint8_t* found_out, void* index_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RTB_ARG(dtype >= RTB_BOOL && dtype <= RTB_F64, "IsMember32: one numeric key column");
  RTB_ARG(index_dtype == RTB_I8 || index_dtype == RTB_I16 || index_dtype == RTB_I32 || index_dtype == RTB_I64, "IsMember32: bad index dtype");
  RTB_ARG(base_index == 0 || base_index == 1, "IsMember32: base_index must be 0 or 1");
  RTB_ARG(na >= 0 && nb >= 0 && nb <= 2147483646ll, "IsMember32: bad sizes");
  if (na == 0) return RTB_OK;
  RTB_ARG(a && found_out && index_out, "IsMember32: null pointer");
  const int isz = dtype_size(dtype);
  long long absent = 0;
  if (!base_index) absent = index_dtype == RTB_I8 ? -128ll : index_dtype == RTB_I16 ? -32768ll : index_dtype == RTB_I32 ? -2147483648ll : (long long)0x8000000000000000ull;
  IsMemberLayout L = ismember_layout(workspace, workspace_bytes, nb);
  RTB_ARG(workspace && L.bytes <= workspace_bytes, "IsMember32: workspace too small");
  RTB_ARG(((uintptr_t)workspace % 256) == 0, "IsMember32: workspace must be 256-byte aligned");
  uint32_t cap = 1024;
  Slot* table = (Slot*)L.gb_ws;  // layout() puts the table first
  if (nb > 0) {
    RTB_ARG(b, "IsMember32: null pointer");
    if (((uintptr_t)b % vec4_align(isz)) != 0) {
      set_error("IsMember32: the key column must be aligned to 4 items");
      return RTB_ERR_UNSUPPORTED;
    }
    const uint8_t* filt = nullptr;
    if (dtype == RTB_F32 || dtype == RTB_F64) {  // NaN never matches: keep it out of the table
      ismember_valid_kernel<<<grid_for(nb, 256, 4), 256, 0, stream>>>(b, dtype, nb, L.valid_b);
      RTB_LAUNCH_CHECK();
      filt = L.valid_b;
    }
    rtb_column col{const_cast<void*>(b), dtype, isz};
    rtb_gb_state st;
    memset(&st, 0, sizeof(st));
    int64_t u = 0;
    int rc = gb_core(&col, 1, nb, filt, nullptr, 0, L.ikey_b, L.ifirst_b, nb, &u, nullptr, nullptr, L.gb_ws, L.gb_bytes, stream, &st, 0);
    if (rc != RTB_OK) return rc;
    cap = st.cap;
    if (cap == 0) {  // every row of b was filtered out: an empty table
      cap = 1024;
      RTB_CUDA(cudaMemsetAsync(table, 0xFF, sizeof(Slot) * ((size_t)cap + 1), stream));
    }
  } else {
    RTB_CUDA(cudaMemsetAsync(table, 0xFF, sizeof(Slot) * ((size_t)cap + 1), stream));
  }
  if (((uintptr_t)a % 16) != 0 || ((uintptr_t)found_out % 16) != 0 || ((uintptr_t)index_out % 16) != 0) {
    set_error("IsMember32: a, found and index buffers must be 16-byte aligned");
    return RTB_ERR_UNSUPPORTED;
  }
  const int grid = grid_for(na, 256, 4, 4);
  switch (isz) {
    case 8: ismember_lookup_kernel<8><<<grid, 256, 0, stream>>>(a, na, table, cap - 1, base_index, index_dtype, absent, found_out, index_out); break;
    case 4: ismember_lookup_kernel<4><<<grid, 256, 0, stream>>>(a, na, table, cap - 1, base_index, index_dtype, absent, found_out, index_out); break;
    case 2: ismember_lookup_kernel<2><<<grid, 256, 0, stream>>>(a, na, table, cap - 1, base_index, index_dtype, absent, found_out, index_out); break;
    default: ismember_lookup_kernel<1><<<grid, 256, 0, stream>>>(a, na, table, cap - 1, base_index, index_dtype, absent, found_out, index_out);
  }
  RTB_LAUNCH_CHECK();
  return RTB_OK;
}
