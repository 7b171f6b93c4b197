// Shared device helpers: sortable keys, warp-cooperative exact selection (radix select +
// in-place compaction) on candidate buffers, small PTX wrappers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ccr {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int kQTile = 128;   // query rows per tensor-core tile (UMMA M)
constexpr int kITile = 256;   // item rows per tensor-core tile (UMMA N)
constexpr int kKBlock = 64;   // K elements per pipeline stage (= one 128-byte swizzle row)
constexpr int kSimtRows = 8;  // query rows per SIMT pass
constexpr int kSimtChunk = 32;  // items per SIMT block iteration

// ---------------------------------------------------------------------------------------
// Sortable keys.  Larger key == better rank (score descending, then item id ascending).
//   dense candidates: 64-bit  [ ord32(score) : ~local_id ]
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 ord32(float f) {
#ifdef __CUDA_ARCH__
  u32 b = __float_as_uint(f);
#else
  union { float f; u32 u; } c; c.f = f; u32 b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord32(u32 o) {
  u32 b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; u32 u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ u64 ord64(double d) {
#ifdef __CUDA_ARCH__
  u64 b = (u64)__double_as_longlong(d);
#else
  union { double d; u64 u; } c; c.d = d; u64 b = c.u;
#endif
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double unord64(u64 o) {
  u64 b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  union { double d; u64 u; } c; c.u = b; return c.d;
#endif
}
__host__ __device__ __forceinline__ u64 make_key(float s, u32 id) {
  return ((u64)ord32(s) << 32) | (u64)(0xFFFFFFFFu - id);
}
__host__ __device__ __forceinline__ float key_score(u64 k) { return unord32((u32)(k >> 32)); }
__host__ __device__ __forceinline__ u32 key_id(u64 k) { return 0xFFFFFFFFu - (u32)k; }

// candidate-buffer capacity for a given k: room for >= k fresh inserts between prunes plus one
// full chunk (32 columns) of slack, multiple of 64
__host__ __device__ __forceinline__ int cand_capacity(int k) { return ((2 * k + 32 + 63) / 64) * 64; }

// ---------------------------------------------------------------------------------------
// mask lookup: is `col` present in the sorted list cols[beg, end) ?
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool mask_contains(const int* __restrict__ cols, long long beg, long long end,
                                              int col) {
  while (beg < end) {
    long long mid = (beg + end) >> 1;
    int c = __ldg(cols + mid);
    if (c == col) return true;
    if (c < col) beg = mid + 1; else end = mid;
  }
  return false;
}

// ---------------------------------------------------------------------------------------
// Warp-cooperative exact selection on a buffer of n unique 64-bit keys in global memory.
// Returns the `kth` largest key (1-based).  hist: 256 x u32 of shared memory private to the
// warp.  MSB-first 8-bit radix select with early exit once the target bin holds one key.
// All 32 lanes must call it with identical arguments.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ u64 warp_select_kth(const u64* buf, int n, int kth, u32* hist) {
  const int lane = threadIdx.x & 31;
  u64 prefix = 0, pmask = 0;
  int need = kth;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      u64 key = buf[i];
      if ((key & pmask) == prefix) atomicAdd(&hist[(u32)(key >> shift) & 255u], 1u);
    }
    __syncwarp();
    // lane L owns digits [8L, 8L+8); higher lanes = higher digits
    u32 c[8];
    u32 lane_sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; lane_sum += c[j]; }
    u32 incl = lane_sum;  // sum over lanes >= lane
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      u32 t = __shfl_down_sync(0xffffffffu, incl, off);
      if (lane + off < 32) incl += t;
    }
    u32 above = incl - lane_sum;
    bool mine = (above < (u32)need) && ((u32)need <= incl);
    u32 d = 0, newneed = 0, binc = 0;
    if (mine) {
      u32 run = above;
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (binc == 0) {
          if (run + c[j] >= (u32)need) { d = lane * 8 + j; newneed = need - run; binc = c[j]; }
          else run += c[j];
        }
      }
    }
    unsigned who = __ballot_sync(0xffffffffu, mine);
    int src = __ffs(who) - 1;  // exactly one lane when kth <= n
    d = __shfl_sync(0xffffffffu, d, src);
    newneed = __shfl_sync(0xffffffffu, newneed, src);
    binc = __shfl_sync(0xffffffffu, binc, src);
    prefix |= (u64)d << shift;
    pmask |= 0xFFull << shift;
    need = (int)newneed;
    __syncwarp();
    if (binc == 1 && shift > 0) {
      // unique key with this prefix: find it
      u64 found = 0;
      for (int i = lane; i < n; i += 32) {
        u64 key = buf[i];
        if ((key & pmask) == prefix) found = key;
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        u64 o = __shfl_xor_sync(0xffffffffu, found, off);
        found = found > o ? found : o;
      }
      return found;
    }
  }
  return prefix;
}

// Keep only keys >= pivot, compacted to the front of buf (stable).  Returns the kept count.
__device__ __forceinline__ int warp_compact_ge(u64* buf, int n, u64 pivot) {
  const int lane = threadIdx.x & 31;
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    int i = i0 + lane;
    u64 key = (i < n) ? buf[i] : 0ull;
    bool keep = (i < n) && (key >= pivot);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) buf[base + __popc(m & ((1u << lane) - 1u))] = key;
    base += __popc(m);
    __syncwarp();
  }
  return base;
}

// prune a buffer down to its top-k; returns pivot (k-th largest key)
__device__ __forceinline__ u64 warp_prune(u64* buf, int n, int k, u32* hist) {
  __syncwarp();
  u64 pivot = warp_select_kth(buf, n, k, hist);
  warp_compact_ge(buf, n, pivot);
  __syncwarp();
  return pivot;
}

// ---------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float bf16_lo(u32 w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(u32 w) { return __uint_as_float(w & 0xffff0000u); }

// error slots written by device-side watchdogs (workspace tail)
struct DeviceStatus {
  int code;     // 0 ok
  int where;    // role / barrier id
  int block;
  int extra;
};

}  // namespace ccr
