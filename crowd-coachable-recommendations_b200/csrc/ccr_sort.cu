// Whole-matrix argsort (rime_lite `_argsort`, src/rime_lite/util/__init__.py:158-184) and the MRR
// first-hit scan (scripts/al_0_rank.py:130-133) -- HBM-bound integer work, CUDA cores only.
//
//   argsort : keys = ~ord64(double(score) [+ prior | := value]) per matrix element, payload = flat
//             index; LSD radix sort, 8 passes of 8 bits, each pass = per-tile digit histogram ->
//             exclusive scan over (digit, tile) -> stable scatter.  Ascending ~ord64 == descending
//             score; stability keeps equal scores in flat-index order (the reference adds unseeded
//             jitter instead).  Algorithmic bytes: 8 passes x 2 x 12 B per element.
//   first-hit rank : one warp per query row scans its ranked ids against the row's sorted list of
//             relevant ids (binary search per id), first lane that hits wins.
#include "ccr_params.cuh"

namespace ccr {

constexpr int kSortThreads = 256;
constexpr int kSortPerThread = 16;
constexpr int kSortTile = kSortThreads * kSortPerThread;  // keys per block and pass

__global__ void __launch_bounds__(256) sort_build_keys_kernel(const float* __restrict__ scores, long long B, long long N,
                                                              long long ld, u64* __restrict__ keys,
                                                              u32* __restrict__ payload) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const long long r = i / N, c = i - r * N;
  keys[i] = ~ord64((double)scores[r * ld + c]);
  payload[i] = (u32)i;
}

// mask entries overwrite their element's key: ADD ranks double(score) + value, SET the value itself
__global__ void __launch_bounds__(256) sort_override_keys_kernel(const float* __restrict__ scores, long long B,
                                                                 long long N, long long ld,
                                                                 const long long* __restrict__ indptr,
                                                                 const int* __restrict__ cols,
                                                                 const double* __restrict__ vals, int mode,
                                                                 u64* __restrict__ keys) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= indptr[B]) return;
  long long lo = 0, hi = B;  // row of entry e: largest r with indptr[r] <= e
  while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (indptr[mid] <= e) lo = mid; else hi = mid; }
  const int col = cols[e];
  if (col < 0 || (long long)col >= N) return;
  const double v = mode == 1 ? vals[e] : (double)scores[lo * ld + col] + vals[e];
  keys[lo * N + col] = ~ord64(v);
}

// per-tile digit counts, stored digit-major: hist[d * n_tiles + tile]
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const u64* __restrict__ keys, long long n, int shift,
                                                                 u32* __restrict__ hist, int n_tiles) {
  __shared__ u32 s_cnt[256];
  const int tid = threadIdx.x;
  s_cnt[tid] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll
  for (int j = 0; j < kSortPerThread; ++j) {
    const long long i = base + j * kSortThreads + tid;
    if (i < n) atomicAdd(&s_cnt[(u32)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(long long)tid * n_tiles + blockIdx.x] = s_cnt[tid];
}

// exclusive scan of the 256 * n_tiles counters in place (one block; the array is a few hundred KB)
__global__ void __launch_bounds__(1024) sort_scan_kernel(u32* __restrict__ hist, long long m) {
  __shared__ u32 s_warp[32];
  __shared__ u32 s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (long long i0 = 0; i0 < m; i0 += 1024) {
    const long long i = i0 + tid;
    const u32 v = i < m ? hist[i] : 0u;
    u32 incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    u32 base = s_carry;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    if (i < m) hist[i] = base + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = base + incl;
    __syncthreads();
  }
}

// Stable scatter.  To stay coalesced AND stable the tile is walked in 16 rounds of 256 consecutive keys;
// inside a round warp w holds keys [32 w, 32 w + 32).  Stable rank of a key = keys of the same digit in
// earlier rounds (accumulated into s_base) + in earlier warps of this round + in lower lanes of its warp.
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const u64* __restrict__ keys_in,
                                                                    const u32* __restrict__ pay_in, long long n,
                                                                    int shift, const u32* __restrict__ hist,
                                                                    int n_tiles, u64* __restrict__ keys_out,
                                                                    u32* __restrict__ pay_out) {
  __shared__ u32 s_base[256];          // global offset of the next key of each digit from this tile
  __shared__ u32 s_wcnt[8][256];       // per-warp digit counts of the current round
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  s_base[tid] = hist[(long long)tid * n_tiles + blockIdx.x];
  const long long base = (long long)blockIdx.x * kSortTile;
  for (int j = 0; j < kSortPerThread; ++j) {
    for (int d = lane; d < 256; d += 32) s_wcnt[warp][d] = 0;
    __syncthreads();
    const long long i = base + j * kSortThreads + tid;
    const bool live = i < n;
    const u64 key = live ? keys_in[i] : 0ull;
    const u32 pay = live ? pay_in[i] : 0u;
    const u32 d = (u32)(key >> shift) & 255u;
    // lanes of this warp with the same digit (dead lanes form their own group)
    const unsigned peers = __match_any_sync(0xffffffffu, live ? d : 0xFFFFFFFFu);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (live && rank_in_warp == 0) s_wcnt[warp][d] = __popc(peers);
    __syncthreads();
    if (live) {
      u32 before = 0;
      for (int w = 0; w < warp; ++w) before += s_wcnt[w][d];
      const u32 pos = s_base[d] + before + (u32)rank_in_warp;
      keys_out[pos] = key;
      pay_out[pos] = pay;
    }
    __syncthreads();
    {  // advance the digit bases by this round's totals
      u32 tot = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += s_wcnt[w][tid];
      s_base[tid] += tot;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sort_unravel_kernel(const u32* __restrict__ payload, long long n, long long N,
                                                           long long* __restrict__ out_rows,
                                                           long long* __restrict__ out_cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long flat = payload[i];
  out_rows[i] = flat / N;
  out_cols[i] = flat - (flat / N) * N;
}

size_t argsort_workspace_bytes(long long n) {
  const long long tiles = (n + kSortTile - 1) / kSortTile;
  size_t off = 0;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  off = up(off + (size_t)n * 8);   // keys a
  off = up(off + (size_t)n * 8);   // keys b
  off = up(off + (size_t)n * 4);   // payload a
  off = up(off + (size_t)n * 4);   // payload b
  off = up(off + (size_t)256 * (size_t)(tiles > 0 ? tiles : 1) * 4);
  return off;
}

int launch_argsort(const float* scores, long long B, long long N, long long ld, const long long* indptr, const int* cols,
                   const double* vals, long long nnz, int mode, long long* out_rows, long long* out_cols, void* ws,
                   cudaStream_t st) {
  const long long n = B * N;
  if (n <= 0) return 0;
  const int tiles = (int)((n + kSortTile - 1) / kSortTile);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  unsigned char* w = (unsigned char*)ws;
  size_t off = 0;
  u64* ka = (u64*)(w + off); off = up(off + (size_t)n * 8);
  u64* kb = (u64*)(w + off); off = up(off + (size_t)n * 8);
  u32* pa = (u32*)(w + off); off = up(off + (size_t)n * 4);
  u32* pb = (u32*)(w + off); off = up(off + (size_t)n * 4);
  u32* hist = (u32*)(w + off);
  const unsigned gb = (unsigned)((n + 255) / 256);
  sort_build_keys_kernel<<<gb, 256, 0, st>>>(scores, B, N, ld, ka, pa);
  if (nnz > 0 && indptr)
    sort_override_keys_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(scores, B, N, ld, indptr, cols, vals, mode, ka);
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = pass * 8;
    sort_hist_kernel<<<tiles, kSortThreads, 0, st>>>(ka, n, shift, hist, tiles);
    sort_scan_kernel<<<1, 1024, 0, st>>>(hist, (long long)256 * tiles);
    sort_scatter_kernel<<<tiles, kSortThreads, 0, st>>>(ka, pa, n, shift, hist, tiles, kb, pb);
    u64* tk = ka; ka = kb; kb = tk;
    u32* tp = pa; pa = pb; pb = tp;
  }
  sort_unravel_kernel<<<gb, 256, 0, st>>>(pa, n, N, out_rows, out_cols);
  return (int)cudaGetLastError();
}

// =======================================================================================
// MRR core: rank (1-based) of the first relevant id in each row's ranked list, 0 = none.
//   ids [B, k] int64 (ranked, best first; < 0 = padding), rel CSR: rel_indptr int64[B+1], rel_ids
//   int64 sorted ascending inside a row.
// =======================================================================================
__global__ void __launch_bounds__(256) first_hit_rank_kernel(const long long* __restrict__ ids, long long B, int k,
                                                             const long long* __restrict__ rel_indptr,
                                                             const long long* __restrict__ rel_ids,
                                                             int* __restrict__ out_rank) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const long long beg = rel_indptr[row], end = rel_indptr[row + 1];
  int rank = 0;
  if (end > beg) {
    for (int j0 = 0; j0 < k && rank == 0; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < k) {
        const long long id = ids[row * k + j];
        long long lo = beg, hi = end;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; const long long v = rel_ids[mid]; if (v == id) { hit = true; break; } if (v < id) lo = mid + 1; else hi = mid; }
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) rank = j0 + __ffs(m);
    }
  }
  if (lane == 0) out_rank[row] = rank;
}

int launch_first_hit_rank(const long long* ids, long long B, int k, const long long* rel_indptr, const long long* rel_ids,
                          int* out_rank, cudaStream_t st) {
  if (B <= 0) return 0;
  first_hit_rank_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(ids, B, k, rel_indptr, rel_ids, out_rank);
  return (int)cudaGetLastError();
}

}  // namespace ccr
