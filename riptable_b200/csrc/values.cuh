// Raw-bit value helpers shared by the reduction, pick and cumsum kernels: a value travels as its zero-extended
// bit pattern (`raw`) plus the rtb_dtype code; sentinels follow riptable/rt_enum.py:88-143 (INVALID_DICT).
#pragma once
#include "common.cuh"

namespace rtb {

// ---- raw value access ---------------------------------------------------------------------------
__device__ __forceinline__ void load_raw4(const void* p, int sz, int64_t r, unsigned long long v[4]) {
  if (sz == 8) {
    uint4 a = ld_stream_u4((const uint8_t*)p + r * 8), b = ld_stream_u4((const uint8_t*)p + r * 8 + 16);
    v[0] = ((unsigned long long)a.y << 32) | a.x; v[1] = ((unsigned long long)a.w << 32) | a.z;
    v[2] = ((unsigned long long)b.y << 32) | b.x; v[3] = ((unsigned long long)b.w << 32) | b.z;
  } else if (sz == 4) {
    uint4 a = ld_stream_u4((const uint8_t*)p + r * 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else if (sz == 2) {
    uint2 a = ld_stream_u2((const uint8_t*)p + r * 2);
    v[0] = a.x & 0xffff; v[1] = a.x >> 16; v[2] = a.y & 0xffff; v[3] = a.y >> 16;
  } else {
    uint32_t a = ld_stream_u32((const uint8_t*)p + r);
    v[0] = a & 0xff; v[1] = (a >> 8) & 0xff; v[2] = (a >> 16) & 0xff; v[3] = a >> 24;
  }
}
// two consecutive items at row r (r even, base 16-byte aligned): one coalesced vector load
__device__ __forceinline__ void load_raw2(const void* p, int sz, int64_t r, unsigned long long v[2]) {
  if (sz == 8) {
    uint4 a = ld_stream_u4((const uint8_t*)p + r * 8);
    v[0] = ((unsigned long long)a.y << 32) | a.x; v[1] = ((unsigned long long)a.w << 32) | a.z;
  } else if (sz == 4) {
    uint2 a = ld_stream_u2((const uint8_t*)p + r * 4);
    v[0] = a.x; v[1] = a.y;
  } else if (sz == 2) {
    uint32_t a = ld_stream_u32((const uint8_t*)p + r * 2);
    v[0] = a & 0xffff; v[1] = a >> 16;
  } else {
    uint32_t a = *(const uint16_t*)((const uint8_t*)p + r);
    v[0] = a & 0xff; v[1] = a >> 8;
  }
}
__device__ __forceinline__ unsigned long long load_raw1(const void* p, int sz, int64_t r) {
  if (sz == 8) return ((const unsigned long long*)p)[r];
  if (sz == 4) return ((const uint32_t*)p)[r];
  if (sz == 2) return ((const uint16_t*)p)[r];
  return ((const uint8_t*)p)[r];
}
__device__ __forceinline__ int dsize(int dt) {
  return dt <= RTB_U8 ? 1 : dt <= RTB_U16 ? 2 : (dt <= RTB_U32 || dt == RTB_F32) ? 4 : 8;
}

__device__ __forceinline__ bool raw_invalid(unsigned long long raw, int dt) {
  switch (dt) {
    case RTB_F64: return (raw & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull;
    case RTB_F32: return ((uint32_t)raw & 0x7FFFFFFFu) > 0x7F800000u;
    case RTB_I8: return raw == 0x80ull;
    case RTB_U8: return raw == 0xFFull;
    case RTB_I16: return raw == 0x8000ull;
    case RTB_U16: return raw == 0xFFFFull;
    case RTB_I32: return raw == 0x80000000ull;
    case RTB_U32: return raw == 0xFFFFFFFFull;
    case RTB_I64: return raw == 0x8000000000000000ull;
    case RTB_U64: return raw == ~0ull;
    default: return false;  // bool has no invalid
  }
}
__device__ __forceinline__ long long raw_to_i64(unsigned long long raw, int dt) {
  switch (dt) {
    case RTB_BOOL: return raw != 0;
    case RTB_I8: return (int8_t)raw;
    case RTB_I16: return (int16_t)raw;
    case RTB_I32: return (int32_t)raw;
    default: return (long long)raw;  // unsigned types are already zero-extended; 64-bit as is
  }
}
__device__ __forceinline__ double raw_to_f64(unsigned long long raw, int dt) {
  switch (dt) {
    case RTB_F64: return __longlong_as_double((long long)raw);
    case RTB_F32: return (double)__uint_as_float((uint32_t)raw);
    case RTB_U64: return (double)raw;
    case RTB_U8: case RTB_U16: case RTB_U32: return (double)(long long)raw;
    default: return (double)raw_to_i64(raw, dt);
  }
}
// Order-preserving code with both ends of the uint64 range left free for the "empty" identity.
__device__ __forceinline__ unsigned long long raw_to_code(unsigned long long raw, int dt, int ismax) {
  switch (dt) {
    case RTB_F64: return (raw >> 63) ? ~raw : (raw | 0x8000000000000000ull);
    case RTB_F32: {
      uint32_t u = (uint32_t)raw;
      u = (u >> 31) ? ~u : (u | 0x80000000u);
      return (unsigned long long)u + (1ull << 32);
    }
    case RTB_I64: {
      unsigned long long c = raw ^ 0x8000000000000000ull;  // sentinel (c == 0) never reaches here
      return ismax ? c : c - 1;
    }
    case RTB_U64: return ismax ? raw + 1 : raw;  // sentinel (~0) never reaches here
    case RTB_BOOL: return (raw != 0) + 1ull;
    case RTB_U8: case RTB_U16: case RTB_U32: return raw + 1;
    default: return (unsigned long long)raw_to_i64(raw, dt) ^ 0x8000000000000000ull;
  }
}
__device__ __forceinline__ unsigned long long code_to_raw(unsigned long long c, int dt, int ismax) {
  switch (dt) {
    case RTB_F64: return (c >> 63) ? (c & 0x7FFFFFFFFFFFFFFFull) : ~c;
    case RTB_F32: {
      uint32_t u = (uint32_t)(c - (1ull << 32));
      return (u >> 31) ? (u & 0x7FFFFFFFu) : (uint32_t)~u;
    }
    case RTB_I64: return (ismax ? c : c + 1) ^ 0x8000000000000000ull;
    case RTB_U64: return ismax ? c - 1 : c;
    case RTB_BOOL: case RTB_U8: case RTB_U16: case RTB_U32: return c - 1;
    default: return c ^ 0x8000000000000000ull;
  }
}
__device__ __forceinline__ unsigned long long invalid_raw(int dt) {
  switch (dt) {
    case RTB_F64: return 0x7FF8000000000000ull;
    case RTB_F32: return 0x7FC00000ull;
    case RTB_I8: return 0x80ull;
    case RTB_U8: return 0xFFull;
    case RTB_I16: return 0x8000ull;
    case RTB_U16: return 0xFFFFull;
    case RTB_I32: return 0x80000000ull;
    case RTB_U32: return 0xFFFFFFFFull;
    case RTB_I64: return 0x8000000000000000ull;
    case RTB_U64: return ~0ull;
    default: return 0;
  }
}
__device__ __forceinline__ void store_raw(void* p, int dt, int64_t i, unsigned long long raw) {
  switch (dsize(dt)) {
    case 1: ((uint8_t*)p)[i] = (uint8_t)raw; break;
    case 2: ((uint16_t*)p)[i] = (uint16_t)raw; break;
    case 4: ((uint32_t*)p)[i] = (uint32_t)raw; break;
    default: ((unsigned long long*)p)[i] = raw;
  }
}

}  // namespace rtb
