// Shared device/host helpers for the riptable_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/riptable_b200.h"

namespace rtb {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern uint64_t g_launches;
extern int g_debug_sync;  // RTB_DEBUG_SYNC=1: synchronise the device after every launch (debugging aid)
inline void count_launch(int n = 1) { g_launches += (uint64_t)n; }

#define RTB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      rtb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return RTB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define RTB_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    rtb::count_launch();                                                                    \
    if (rtb::g_debug_sync) cudaDeviceSynchronize();                                         \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      rtb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return RTB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define RTB_ARG(cond, ...)            \
  do {                                \
    if (!(cond)) {                    \
      rtb::set_error(__VA_ARGS__);    \
      return RTB_ERR_ARG;             \
    }                                 \
  } while (0)

// ---- optional per-kernel timing (CUDA events on the launching stream; bench.py's roofline leg) --------
enum ProfId { PROF_GB_ROUND = 0, PROF_ACCUM = 1, PROF_RADIX_HIST = 2, PROF_RADIX_SCATTER = 3, PROF_SEGSCAN = 4, PROF_GATHER = 5, PROF_COUNT = 8 };
extern int g_prof_on;
void prof_begin(int id, cudaStream_t s);
void prof_end(int id, cudaStream_t s, uint64_t units);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  uint8_t* base;
  size_t size, off;
  Arena(void* p, size_t n) : base((uint8_t*)p), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = (T*)(base ? base + off : nullptr);
    off += count * sizeof(T);
    return p;
  }
  bool ok() const { return off <= size; }
};

inline int dtype_size(int dt) {
  switch (dt) {
    case RTB_BOOL: case RTB_I8: case RTB_U8: return 1;
    case RTB_I16: case RTB_U16: return 2;
    case RTB_I32: case RTB_U32: case RTB_F32: return 4;
    case RTB_I64: case RTB_U64: case RTB_F64: return 8;
    default: return 0;
  }
}
// Alignment a column's base pointer needs for the kernels' 4-row vector loads: 4 items, but never more than the widest
// load there is (16 bytes) -- a 16-byte aligned slice of an int64 / float64 tensor is fine.
inline size_t vec4_align(int itemsize) { return itemsize >= 4 ? 16 : (size_t)(4 * itemsize); }
inline bool is_ikey_dtype(int dt) { return dt == RTB_I8 || dt == RTB_I16 || dt == RTB_I32 || dt == RTB_I64; }

inline int grid_for(int64_t work_items, int threads, int items_per_thread, int max_blocks_per_sm = 16) {
  int64_t per_block = (int64_t)threads * items_per_thread;
  int64_t b = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)kSMs * max_blocks_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- streaming loads / stores: keep the one-pass row streams out of the way of L2-resident tables
// RTB_STREAM_CS=1 (measurement build): the `.cs` ("cache streaming, likely accessed once": evict-first in L1 and L2) forms
// instead of `.nc.L1::no_allocate`, to see whether the key / value / iKey streams can be kept from evicting the hash table.
#ifndef RTB_STREAM_CS
#define RTB_STREAM_CS 0
#endif
#if RTB_STREAM_CS
#define RTB_LD_STREAM "ld.global.cs"
#define RTB_ST_STREAM "st.global.cs"
#else
#define RTB_LD_STREAM "ld.global.nc.L1::no_allocate"
#define RTB_ST_STREAM "st.global.L1::no_allocate"
#endif
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile(RTB_LD_STREAM ".v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
  uint2 r;
  asm volatile(RTB_LD_STREAM ".v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t r;
  asm volatile(RTB_LD_STREAM ".u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile(RTB_ST_STREAM ".v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(void* p, uint2 v) {
  asm volatile(RTB_ST_STREAM ".v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// ---- iKey loads (int8/16/32/64 -> int64), 4 consecutive rows at r (r % 4 == 0, r+3 < n)
__device__ __forceinline__ int64_t load_ikey(const void* p, int kdt, int64_t r) {
  switch (kdt) {
    case RTB_I8: return ((const int8_t*)p)[r];
    case RTB_I16: return ((const int16_t*)p)[r];
    case RTB_I32: return ((const int32_t*)p)[r];
    default: return ((const int64_t*)p)[r];
  }
}
__device__ __forceinline__ void load_ikey4(const void* p, int kdt, int64_t r, int64_t k[4]) {
  switch (kdt) {
    case RTB_I8: {
      uint32_t w = ld_stream_u32((const int8_t*)p + r);
      k[0] = (int8_t)(w & 0xff); k[1] = (int8_t)((w >> 8) & 0xff); k[2] = (int8_t)((w >> 16) & 0xff); k[3] = (int8_t)(w >> 24);
      break;
    }
    case RTB_I16: {
      uint2 w = ld_stream_u2((const int16_t*)p + r);
      k[0] = (int16_t)(w.x & 0xffff); k[1] = (int16_t)(w.x >> 16); k[2] = (int16_t)(w.y & 0xffff); k[3] = (int16_t)(w.y >> 16);
      break;
    }
    case RTB_I32: {
      uint4 w = ld_stream_u4((const int32_t*)p + r);
      k[0] = (int32_t)w.x; k[1] = (int32_t)w.y; k[2] = (int32_t)w.z; k[3] = (int32_t)w.w;
      break;
    }
    default: {
      uint4 a = ld_stream_u4((const int64_t*)p + r), b = ld_stream_u4((const int64_t*)p + r + 2);
      k[0] = (int64_t)(((uint64_t)a.y << 32) | a.x); k[1] = (int64_t)(((uint64_t)a.w << 32) | a.z);
      k[2] = (int64_t)(((uint64_t)b.y << 32) | b.x); k[3] = (int64_t)(((uint64_t)b.w << 32) | b.z);
    }
  }
}
// two consecutive iKeys at row r (r even)
__device__ __forceinline__ void load_ikey2(const void* p, int kdt, int64_t r, int64_t k[2]) {
  switch (kdt) {
    case RTB_I8: {
      uint32_t w = *(const uint16_t*)((const int8_t*)p + r);
      k[0] = (int8_t)(w & 0xff); k[1] = (int8_t)(w >> 8);
      break;
    }
    case RTB_I16: {
      uint32_t w = ld_stream_u32((const int16_t*)p + r);
      k[0] = (int16_t)(w & 0xffff); k[1] = (int16_t)(w >> 16);
      break;
    }
    case RTB_I32: {
      uint2 w = ld_stream_u2((const int32_t*)p + r);
      k[0] = (int32_t)w.x; k[1] = (int32_t)w.y;
      break;
    }
    default: {
      uint4 a = ld_stream_u4((const int64_t*)p + r);
      k[0] = (int64_t)(((uint64_t)a.y << 32) | a.x); k[1] = (int64_t)(((uint64_t)a.w << 32) | a.z);
    }
  }
}
__device__ __forceinline__ void store_ikey(void* p, int kdt, int64_t r, int64_t v) {
  switch (kdt) {
    case RTB_I8: ((int8_t*)p)[r] = (int8_t)v; break;
    case RTB_I16: ((int16_t*)p)[r] = (int16_t)v; break;
    case RTB_I32: ((int32_t*)p)[r] = (int32_t)v; break;
    default: ((int64_t*)p)[r] = v;
  }
}

// ---- device-wide exclusive scan of uint32 (three launches; sums must fit in uint32)
enum ScanXform { SCAN_IDENTITY = 0, SCAN_POPC = 1 };
size_t scan_u32_workspace_elems(int64_t n);  // uint32 elements of scratch
// out[i] = sum_{j<i} f(in[j]); if total_dev != nullptr, *total_dev = sum of all.  in may alias out.
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, ScanXform xf, uint32_t* scratch,
                       uint32_t* total_dev, cudaStream_t stream);

}  // namespace rtb
