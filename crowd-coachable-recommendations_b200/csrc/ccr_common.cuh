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

// candidate-buffer capacity for a given k: room for >= k fresh inserts between prunes plus
// `slack` columns that may be appended before the next prune opportunity (32 for the SIMT
// kernel: one block iteration; 128 for the tensor-core kernel: one half tile), multiple of 64
__host__ __device__ __forceinline__ int cand_capacity(int k, int slack) { return ((2 * k + slack + 63) / 64) * 64; }

// smallest float strictly greater than x (x finite): "s > x"  <=>  "s >= next_up(x)"
__device__ __forceinline__ float next_up(float x) {
  if (!(x == x) || x == INFINITY) return x;
  if (x == 0.f) return __uint_as_float(1u);
  u32 b = __float_as_uint(x);
  return __uint_as_float(x > 0.f ? b + 1u : b - 1u);
}

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
// explicit shared-memory accessors (32-bit shared addresses): keeps the selection code on
// LDS / STS / ATOMS instead of generic LD / ST / ATOM when pointers lose their address space
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sm_red_inc(u32 a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
__device__ __forceinline__ u32 sm_ld32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sm_st32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ u64 sm_ld64(u32 a) { u64 v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sm_st64(u32 a, u64 v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

struct GlobalKeys {
  u64* p;
  __device__ __forceinline__ u64 get(int i) const { return p[i]; }
  __device__ __forceinline__ void set(int i, u64 v) const { p[i] = v; }
};
struct SharedKeys {
  u32 s;
  __device__ __forceinline__ u64 get(int i) const { return sm_ld64(s + (u32)i * 8u); }
  __device__ __forceinline__ void set(int i, u64 v) const { sm_st64(s + (u32)i * 8u, v); }
};

// ---------------------------------------------------------------------------------------
// Warp-cooperative exact selection on n unique 64-bit keys.  Returns the `kth` largest key
// (1-based).  hist_s: shared address of 256 x u32 private to the warp.  MSB-first 8-bit radix
// select with early exit once the target bin holds one key.  All 32 lanes call it with
// identical arguments.
// ---------------------------------------------------------------------------------------
template <class Keys>
__device__ __forceinline__ u64 warp_select_kth(Keys keys, int n, int kth, u32 hist_s) {
  const int lane = threadIdx.x & 31;
  u64 prefix = 0, pmask = 0;
  int need = kth;
  for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sm_st32(hist_s + (u32)(lane * 8 + j) * 4u, 0u);
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      u64 key = keys.get(i);
      if ((key & pmask) == prefix) sm_red_inc(hist_s + ((u32)(key >> shift) & 255u) * 4u);
    }
    __syncwarp();
    // lane L owns digits [8L, 8L+8); higher lanes = higher digits
    u32 c[8];
    u32 lane_sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = sm_ld32(hist_s + (u32)(lane * 8 + j) * 4u); lane_sum += c[j]; }
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
        u64 key = keys.get(i);
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

// Keep only keys >= pivot: src[0,n) -> dst[0, kept), stable.  dst may alias src (in place).
template <class Src, class Dst>
__device__ __forceinline__ int warp_compact_ge(Src src, Dst dst, int n, u64 pivot) {
  const int lane = threadIdx.x & 31;
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    int i = i0 + lane;
    u64 key = (i < n) ? src.get(i) : 0ull;
    bool keep = (i < n) && (key >= pivot);
    unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) dst.set(base + __popc(m & ((1u << lane) - 1u)), key);
    base += __popc(m);
    __syncwarp();
  }
  return base;
}

// Prune a candidate buffer in global memory down to its exact top-k; returns the pivot (k-th
// largest key).  With stage_s != 0 the keys are first copied to that shared-memory staging area
// (>= n keys) so the radix passes run at shared-memory latency.
__device__ __forceinline__ u64 warp_prune(u64* buf, int n, int k, u32 hist_s, u32 stage_s) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  u64 pivot;
  if (stage_s) {
    for (int i = lane; i < n; i += 32) sm_st64(stage_s + (u32)i * 8u, buf[i]);
    __syncwarp();
    pivot = warp_select_kth(SharedKeys{stage_s}, n, k, hist_s);
    warp_compact_ge(SharedKeys{stage_s}, GlobalKeys{buf}, n, pivot);
  } else {
    pivot = warp_select_kth(GlobalKeys{buf}, n, k, hist_s);
    warp_compact_ge(GlobalKeys{buf}, GlobalKeys{buf}, n, pivot);
  }
  __syncwarp();
  return pivot;
}

// ---------------------------------------------------------------------------------------
// One-pass approximate prune (the common case).  Keys are staged in shared memory, bucketed by
// the 8 score bits just below the bits all n keys share, and everything at or above the bucket
// where the running count from the top reaches k is kept: between k and k + (that bucket's
// population - 1) survivors, which is all a *threshold* needs (any value with >= k candidates at
// or above it is a valid lower bound of the k-th best).  The same histogram gives, for free, a
// value with >= j candidates at or above it (the stream's contribution to the shared row bound).
// Returns false (nothing modified) when the keys do not spread over buckets (ties) or the
// boundary bucket is too crowded; the caller then falls back to the exact radix select.
//   out_pivot_ord : ord32 score threshold; survivors are exactly the keys with ord >= it
//   out_kept      : survivors, compacted to buf[0, kept)
//   out_j_ord     : ord32 value with at least j keys >= it
// ---------------------------------------------------------------------------------------
// kStaged: keys are first copied to shared memory (stage_s) and the later passes read them there;
// otherwise (buffers larger than the staging area) every pass streams the keys from global memory.
template <bool kStaged>
__device__ __forceinline__ bool warp_prune_hist(u64* buf, int n, int k, int j, int max_keep, u32 hist_s,
                                                u32 stage_s, u32* out_pivot_ord, int* out_kept, u32* out_j_ord) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  u32 mx = 0u, mn = 0xFFFFFFFFu;
  for (int i = lane; i < n; i += 32) {
    const u64 key = buf[i];
    if (kStaged) sm_st64(stage_s + (u32)i * 8u, key);
    const u32 o = (u32)(key >> 32);
    mx = o > mx ? o : mx;
    mn = o < mn ? o : mn;
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const u32 a = __shfl_xor_sync(0xffffffffu, mx, off), b = __shfl_xor_sync(0xffffffffu, mn, off);
    mx = a > mx ? a : mx;
    mn = b < mn ? b : mn;
  }
  const u32 d = mx ^ mn;
  if (d == 0u) return false;
  const int hb = 31 - __clz(d);
  const int shift = hb > 7 ? hb - 7 : 0;
  const u32 base = mn >> shift;  // digit = (ord >> shift) - base  in [0, 255]
#pragma unroll
  for (int t = 0; t < 8; ++t) sm_st32(hist_s + (u32)(lane * 8 + t) * 4u, 0u);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const u32 o = kStaged ? (u32)(sm_ld64(stage_s + (u32)i * 8u) >> 32) : (u32)(buf[i] >> 32);
    sm_red_inc(hist_s + ((o >> shift) - base) * 4u);
  }
  __syncwarp();
  u32 c[8], lane_sum = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) { c[t] = sm_ld32(hist_s + (u32)(lane * 8 + t) * 4u); lane_sum += c[t]; }
  u32 incl = lane_sum;  // keys in bins owned by lanes >= lane
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const u32 t = __shfl_down_sync(0xffffffffu, incl, off);
    if (lane + off < 32) incl += t;
  }
  const u32 above = incl - lane_sum;
  // bin where the count from the top reaches `need`, and that cumulative count
  u32 bin_k = 0, cum_k = 0, bin_j = 0;
  {
    u32 run = above;
    bool done_k = !((above < (u32)k) && ((u32)k <= incl)), done_j = !((above < (u32)j) && ((u32)j <= incl));
#pragma unroll
    for (int t = 7; t >= 0; --t) {
      run += c[t];
      if (!done_k && run >= (u32)k) { bin_k = lane * 8 + t; cum_k = run; done_k = true; }
      if (!done_j && run >= (u32)j) { bin_j = lane * 8 + t; done_j = true; }
    }
  }
  const unsigned who_k = __ballot_sync(0xffffffffu, (above < (u32)k) && ((u32)k <= incl));
  const unsigned who_j = __ballot_sync(0xffffffffu, (above < (u32)j) && ((u32)j <= incl));
  const int src_k = __ffs(who_k) - 1, src_j = __ffs(who_j) - 1;
  bin_k = __shfl_sync(0xffffffffu, bin_k, src_k);
  cum_k = __shfl_sync(0xffffffffu, cum_k, src_k);
  bin_j = __shfl_sync(0xffffffffu, bin_j, src_j);
  if ((int)cum_k > max_keep) return false;
  u32 pivot_ord = (base + bin_k) << shift;
  if (pivot_ord < 0x00800000u) pivot_ord = 0x00800000u;  // never below ord32(-FLT_MAX): stays a finite float
  // stable compaction to buf[0, kept); in place it only ever writes at or below the index it read
  int wbase = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const u64 key = (i < n) ? (kStaged ? sm_ld64(stage_s + (u32)i * 8u) : buf[i]) : 0ull;
    const bool keep = (i < n) && ((u32)(key >> 32) >= pivot_ord);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) buf[wbase + __popc(m & ((1u << lane) - 1u))] = key;
    wbase += __popc(m);
    __syncwarp();
  }
  *out_pivot_ord = pivot_ord;
  *out_kept = wbase;
  u32 j_ord = (base + bin_j) << shift;
  if (j_ord < 0x00800000u) j_ord = 0x00800000u;
  *out_j_ord = j_ord;
  return true;
}

// Append one candidate (noinline: keeps the unrolled per-column hit tests small).
static __device__ __noinline__ int cand_insert(float s, u32 col, int cnt, u64 tau_key, u64* buf,
                                        const int* mask_cols, long long mbeg, long long mend) {
  const u64 key = make_key(s, col);
  if (key > tau_key && !(mask_cols && mask_contains(mask_cols, mbeg, mend, (int)col))) {
    buf[cnt] = key;
    return cnt + 1;
  }
  return cnt;
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
