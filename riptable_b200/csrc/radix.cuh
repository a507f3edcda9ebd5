// Stable LSD radix sort of (key, row-id) pairs on 8-bit digits — the engine behind rc.BinCount(pack=True),
// rc.GroupFromBinCount, rc.EmaAll32 (cumsum), rc.MakeiNext and rc.LexSort32.
//
// One digit pass = three launches over fixed contiguous row chunks (one chunk per CTA):
//   radix_chunk_hist : per-chunk digit counts              (reads keys)
//   exclusive scan   : digit-major scan of counts[256][nchunks]
//   radix_scatter    : per tile, warp-level match ranking in row order (stable), tile staged in shared
//                      memory in digit order so the global stores of one digit are contiguous
// Digits whose bits are identical in every key (found by one OR/AND reduction over the keys) are skipped.
#pragma once
#include "common.cuh"

namespace rtb {

constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 16;
constexpr int kRadixTile = kRadixThreads * kRadixItems;  // 4096 rows per tile

struct RadixLayout {
  uint32_t* counts;   // [256][nchunks]
  uint32_t* scan_scratch;
  unsigned long long* bits;  // {or, and}
  int nchunks;
  int64_t chunk_rows;
  size_t bytes;
};

RadixLayout radix_layout(void* ws, size_t wsbytes, int64_t n);
inline size_t radix_workspace_bytes(int64_t n) { return radix_layout(nullptr, 0, n).bytes + 256; }

// Sorts n pairs by key bits [0, key_bits) (key_bits <= 8 * sizeof(K)); digits whose bits do not vary are skipped.
// Buffers ping-pong between (ka, va) and (kb, vb); if va_iota, the values of the input are 0..n-1 and va is never
// read.  *result_in_b tells which pair holds the result.  Synchronises the stream once (digit plan on the host).
template <typename K>
int radix_sort_pairs(K* ka, uint32_t* va, K* kb, uint32_t* vb, int64_t n, int key_bits, bool va_iota, void* ws,
                     size_t wsbytes, cudaStream_t stream, int* result_in_b, int* passes_run = nullptr);

}  // namespace rtb
