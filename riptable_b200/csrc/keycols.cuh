// Multi-column key rows: descriptor + bytewise hash / equality (shared by hash grouping and GroupFromLexSort).
#pragma once
#include "common.cuh"

namespace rtb {

constexpr int kMaxKeyCols = 16;

struct KeyCols {
  const uint8_t* ptr[kMaxKeyCols];
  int32_t itemsize[kMaxKeyCols];
  int32_t chunk[kMaxKeyCols];  // widest load (16/8/4/2/1) dividing both itemsize and the base address
  int32_t ncols;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long w) {
  h ^= w;
  h *= 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  return h;
}

__device__ __forceinline__ unsigned long long hash_row(const KeyCols& kc, int64_t r) {
  unsigned long long h = 0x243F6A8885A308D3ull;
  for (int c = 0; c < kc.ncols; c++) {
    const uint8_t* p = kc.ptr[c] + r * (int64_t)kc.itemsize[c];
    int sz = kc.itemsize[c];
    switch (kc.chunk[c]) {
      case 16:
        for (int o = 0; o < sz; o += 16) {
          uint4 w = *(const uint4*)(p + o);
          h = mix64(h, ((unsigned long long)w.y << 32) | w.x);
          h = mix64(h, ((unsigned long long)w.w << 32) | w.z);
        }
        break;
      case 8:
        for (int o = 0; o < sz; o += 8) h = mix64(h, *(const unsigned long long*)(p + o));
        break;
      case 4:
        for (int o = 0; o < sz; o += 4) h = mix64(h, *(const uint32_t*)(p + o));
        break;
      case 2:
        for (int o = 0; o < sz; o += 2) h = mix64(h, *(const uint16_t*)(p + o));
        break;
      default:
        for (int o = 0; o < sz; o++) h = mix64(h, p[o]);
    }
  }
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return h;
}

__device__ __forceinline__ bool rows_equal(const KeyCols& kc, int64_t r1, int64_t r2) {
  for (int c = 0; c < kc.ncols; c++) {
    int sz = kc.itemsize[c];
    const uint8_t* p = kc.ptr[c] + r1 * (int64_t)sz;
    const uint8_t* q = kc.ptr[c] + r2 * (int64_t)sz;
    switch (kc.chunk[c]) {
      case 16:
        for (int o = 0; o < sz; o += 16) {
          uint4 a = *(const uint4*)(p + o), b = *(const uint4*)(q + o);
          if (a.x != b.x || a.y != b.y || a.z != b.z || a.w != b.w) return false;
        }
        break;
      case 8:
        for (int o = 0; o < sz; o += 8)
          if (*(const unsigned long long*)(p + o) != *(const unsigned long long*)(q + o)) return false;
        break;
      case 4:
        for (int o = 0; o < sz; o += 4)
          if (*(const uint32_t*)(p + o) != *(const uint32_t*)(q + o)) return false;
        break;
      case 2:
        for (int o = 0; o < sz; o += 2)
          if (*(const uint16_t*)(p + o) != *(const uint16_t*)(q + o)) return false;
        break;
      default:
        for (int o = 0; o < sz; o++)
          if (p[o] != q[o]) return false;
    }
  }
  return true;
}


// 0 iff the two key rows are bytewise equal.  No early exit: every load is issued before any result is needed, so
// the gathers of one comparison (and of neighbouring comparisons) overlap.
__device__ __forceinline__ unsigned long long rows_diff(const KeyCols& kc, int64_t r1, int64_t r2) {
  unsigned long long acc = 0;
  for (int c = 0; c < kc.ncols; c++) {
    int sz = kc.itemsize[c];
    const uint8_t* p = kc.ptr[c] + r1 * (int64_t)sz;
    const uint8_t* q = kc.ptr[c] + r2 * (int64_t)sz;
    switch (kc.chunk[c]) {
      case 16:
        for (int o = 0; o < sz; o += 16) {
          uint4 a = *(const uint4*)(p + o), b = *(const uint4*)(q + o);
          acc |= (unsigned long long)((a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w));
        }
        break;
      case 8:
        for (int o = 0; o < sz; o += 8) acc |= *(const unsigned long long*)(p + o) ^ *(const unsigned long long*)(q + o);
        break;
      case 4:
        for (int o = 0; o < sz; o += 4) acc |= (unsigned long long)(*(const uint32_t*)(p + o) ^ *(const uint32_t*)(q + o));
        break;
      case 2:
        for (int o = 0; o < sz; o += 2) acc |= (unsigned long long)(*(const uint16_t*)(p + o) ^ *(const uint16_t*)(q + o));
        break;
      default:
        for (int o = 0; o < sz; o++) acc |= (unsigned long long)(p[o] ^ q[o]);
    }
  }
  return acc;
}

inline int chunk_for(const void* base, int itemsize) {
  for (int c = 16; c > 1; c >>= 1)
    if (itemsize % c == 0 && ((uintptr_t)base % c) == 0) return c;
  return 1;
}

// Fills a KeyCols from the C-ABI column list; returns false if there are too many columns.
inline bool make_keycols(const rtb_column* keys, int nkeys, KeyCols* kc) {
  memset(kc, 0, sizeof(*kc));
  if (nkeys < 1 || nkeys > kMaxKeyCols - 1) return false;
  kc->ncols = nkeys;
  for (int c = 0; c < nkeys; c++) {
    kc->ptr[c] = (const uint8_t*)keys[c].data;
    kc->itemsize[c] = keys[c].itemsize;
    kc->chunk[c] = chunk_for(keys[c].data, keys[c].itemsize);
  }
  return true;
}

}  // namespace rtb
